/*
 * oracle/particles2d.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's particle-laden
 * D2Q9 hot path (cheryli/MGLC, MPI/Micro_particles/fortran/case4/mpi_particle/, "P4" below): MRT fluid with
 * a solid mask, quadratic-interpolated moving-boundary bounce-back on circular particles, momentum-exchange
 * force/torque, spring repulsion, explicit particle kinematics, refill of uncovered nodes, and the 2-/3-deep
 * halo exchanges.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 *
 * PARITY PIN: Fortran+MPI cannot be built here.  The per-cell / per-link / per-particle arithmetic restated
 * here is checked bit for bit against vectors obtained by machine-evaluating the reference's own source text
 * (tests/golden/make_golden_particles.py); copies (streaming, wall bounce-back, exchanges, mask rebuild) by
 * construction tests.  The reference seeds its 64 particle positions with compiler-specific random_number
 * (P4/initial.F90:52-75), so positions are an INPUT here.  Whole run: the program's text is evaluated as a whole on one
 * rank (56 x 72, two particles: initial() and 30 iterations of collision, streaming, bounceback, bounceback_particle with
 * calQ called as written, macro, calForce, updateCenter incl. mask rebuild and refill, then check();
 * make_golden_particles_run.py -> ref_fortran_particles_run.npz) and this file reproduces every population incl. the ghost
 * layers, rho, u, v, the solid mask, rhoAvg and the particle state bit for bit.
 *
 * Layout is the reference's (P4/freeall.F90:13-17): f(0:8,-2:nx+3,-2:ny+3), f_post(0:8,-1:nx+2,-1:ny+2),
 * obst/obstNew(0:nx+1,0:ny+1) integer, rho,u,v,up,vp(nx,ny); column-major.  Left-to-right expressions,
 * x**2.0d0 as x*x, -ffp-contract=off.  One process emulates all ranks; the particle state every rank holds
 * after the reference's Allreduces is kept once.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q9 9
static const int ex[Q9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};      /* P4/commondata.F90:47-48 */
static const int ey[Q9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int rr[Q9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};         /* opposite, :49-50 */

typedef struct p2_params {
    int total_nx, total_ny, N;
    double rho0, rhoSolid, viscosity, tauf, Snu, Sq, gravity, thresholdWall, stiffWall, thresholdParticle,
           stiffParticle, radius0, Pi;
    /* options of the reference's other particle scenario, MPI/Micro_particles/fortran/case1/mpi_complete ("P1"): 0 / 0.0 = P4 */
    double Uwall;        /* top wall moves with +Uwall, bottom wall with -Uwall along x (P1/fluid.F90:123-171)          */
    double Uframe;       /* U0 of the `movingFrame` build: the walls' velocities are seen from a frame moving with U0    */
    int bb_linear;       /* 1: linear-interpolated bounce-back on the particles (`#ifdef linear`, P1/particle_bounceback.F90:66-76) */
    int moving_walls;    /* 1: apply the wall terms above                                                                 */
} p2_params;

typedef struct p2_rank {
    int nx, ny, coords[2], i_start, j_start;
    int nbr_left, nbr_right, nbr_top, nbr_bottom, cnr_tl, cnr_tr, cnr_bl, cnr_br;
    double *f, *f_post, *rho, *u, *v, *up, *vp;
    int *obst, *obstNew;
} p2_rank;

typedef struct p2_world {
    p2_params p;
    int dims[2], np, itc, error_flag;
    double rhoAvg, errorU, omega[Q9];
    /* particle state, size N */
    double *xCenter, *yCenter, *xCenterOld, *yCenterOld, *Uc, *Vc, *UcOld, *VcOld, *rOmega, *rOmegaOld, *radius;
    double *wallTotalForceX, *wallTotalForceY, *totalTorque;
    p2_rank *r;
} p2_world;

#define FI(R, a, i, j) ((R)->f[(a) + Q9 * ((size_t)((i) + 2) + (size_t)((R)->nx + 6) * (size_t)((j) + 2))])
#define FP(R, a, i, j) ((R)->f_post[(a) + Q9 * ((size_t)((i) + 1) + (size_t)((R)->nx + 4) * (size_t)((j) + 1))])
#define OB(R, A, i, j) ((R)->A[(size_t)(i) + (size_t)((R)->nx + 2) * (size_t)(j)])
#define S2(R, A, i, j) ((R)->A[(size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1)])

/* defaults of module commondata, P4/commondata.F90:3-62 */
void p2_default_params(p2_params *p) {
    const double l0 = 1.0 / 100.0, t0 = 5.0 / 10000.0;
    p->total_nx = 201; p->total_ny = 801; p->N = 64;
    p->rho0 = 1.0; p->rhoSolid = 1.01; p->viscosity = 0.05;
    p->radius0 = 20.0 / 2.0;
    p->thresholdWall = 6.0; p->stiffWall = 0.02; p->thresholdParticle = 6.0; p->stiffParticle = 0.08;
    p->gravity = 980.0 * (t0 * t0) / l0;
    p->Pi = 4.0 * atan(1.0);
    p->tauf = 3.0 * p->viscosity + 0.5;
    p->Snu = 1.0 / p->tauf;
    p->Sq = 8.0 * (2.0 * p->tauf - 1.0) / (8.0 * p->tauf - 1.0);
    p->Uwall = 0.0; p->Uframe = 0.0; p->bb_linear = 0; p->moving_walls = 0;
}

/* MPI_Dims_create_2d, P4/mpi_starts.F90:160-180: the factorisation with the smallest halo message, in
 * default-real (single precision) arithmetic, first minimum wins */
void p2_dims_create(int np, int total_nx, int total_ny, int dims[2]) {
    float diff = (float)(total_nx + total_ny) * (float)np;
    dims[0] = np; dims[1] = 1;
    for (int i = 1; i <= np; ++i)
        for (int j = 1; j <= np; ++j)
            if (i * j == np) {
                float message = (float)(i - 1) * (float)total_ny + (float)(j - 1) * (float)total_nx;
                if (message < diff) { diff = message; dims[0] = i; dims[1] = j; }
            }
}

/* decompose_1d, P4/mpi_starts.F90:95-112 */
static void decompose_1d(int total_n, int *local_n, int rank, int np, int *start) {
    *local_n = total_n / np;
    if (rank < total_n % np) *local_n += 1;
    if (*local_n > total_n / np) *start = *local_n * rank;
    else *start = *local_n * rank + total_n % np;
}
static int cart_rank2(const int dims[2], int c0, int c1) {
    if (c0 < 0 || c0 >= dims[0] || c1 < 0 || c1 >= dims[1]) return -1;
    return c0 * dims[1] + c1;
}

p2_world *p2_world_create(const p2_params *p, int np, const int *dims_or_null) {
    p2_world *w = calloc(1, sizeof *w);
    w->p = *p; w->np = np;
    if (dims_or_null && dims_or_null[0] > 0) { w->dims[0] = dims_or_null[0]; w->dims[1] = dims_or_null[1]; }
    else p2_dims_create(np, p->total_nx, p->total_ny, w->dims);
    const int N = p->N;
    double **pa[] = {&w->xCenter, &w->yCenter, &w->xCenterOld, &w->yCenterOld, &w->Uc, &w->Vc, &w->UcOld, &w->VcOld,
                     &w->rOmega, &w->rOmegaOld, &w->radius, &w->wallTotalForceX, &w->wallTotalForceY, &w->totalTorque};
    for (size_t q = 0; q < sizeof pa / sizeof *pa; ++q) *pa[q] = calloc((size_t)N, 8);
    w->r = calloc((size_t)np, sizeof(p2_rank));
    for (int c0 = 0; c0 < w->dims[0]; ++c0)
    for (int c1 = 0; c1 < w->dims[1]; ++c1) {
        p2_rank *R = &w->r[cart_rank2(w->dims, c0, c1)];
        R->coords[0] = c0; R->coords[1] = c1;
        decompose_1d(p->total_nx, &R->nx, c0, w->dims[0], &R->i_start);
        decompose_1d(p->total_ny, &R->ny, c1, w->dims[1], &R->j_start);
        R->nbr_left = cart_rank2(w->dims, c0 - 1, c1); R->nbr_right = cart_rank2(w->dims, c0 + 1, c1);   /* :34-35 */
        R->nbr_bottom = cart_rank2(w->dims, c0, c1 - 1); R->nbr_top = cart_rank2(w->dims, c0, c1 + 1);
        R->cnr_tr = cart_rank2(w->dims, c0 + 1, c1 + 1); R->cnr_br = cart_rank2(w->dims, c0 + 1, c1 - 1);
        R->cnr_tl = cart_rank2(w->dims, c0 - 1, c1 + 1); R->cnr_bl = cart_rank2(w->dims, c0 - 1, c1 - 1);
        size_t n = (size_t)R->nx * R->ny;
        R->f = calloc(Q9 * (size_t)(R->nx + 6) * (R->ny + 6), 8);
        R->f_post = calloc(Q9 * (size_t)(R->nx + 4) * (R->ny + 4), 8);
        R->rho = calloc(n, 8); R->u = calloc(n, 8); R->v = calloc(n, 8); R->up = calloc(n, 8); R->vp = calloc(n, 8);
        R->obst = calloc((size_t)(R->nx + 2) * (R->ny + 2), sizeof(int));
        R->obstNew = calloc((size_t)(R->nx + 2) * (R->ny + 2), sizeof(int));
    }
    return w;
}

void p2_world_destroy(p2_world *w) {
    if (!w) return;
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        free(R->f); free(R->f_post); free(R->rho); free(R->u); free(R->v); free(R->up); free(R->vp); free(R->obst); free(R->obstNew);
    }
    double *pa[] = {w->xCenter, w->yCenter, w->xCenterOld, w->yCenterOld, w->Uc, w->Vc, w->UcOld, w->VcOld, w->rOmega,
                    w->rOmegaOld, w->radius, w->wallTotalForceX, w->wallTotalForceY, w->totalTorque};
    for (size_t q = 0; q < sizeof pa / sizeof *pa; ++q) free(pa[q]);
    free(w->r); free(w);
}

/* which: 0 f, 1 f_post, 2 rho, 3 u, 4 v, 5 up, 6 vp */
double *p2_rank_ptr(p2_world *w, int r, int which) {
    p2_rank *R = &w->r[r];
    double *p[] = {R->f, R->f_post, R->rho, R->u, R->v, R->up, R->vp};
    return p[which];
}
int *p2_rank_obst(p2_world *w, int r, int which_new) { return which_new ? w->r[r].obstNew : w->r[r].obst; }
void p2_rank_info(p2_world *w, int r, int *out /*[14]*/) {
    p2_rank *R = &w->r[r];
    int v[14] = {R->nx, R->ny, R->coords[0], R->coords[1], R->i_start, R->j_start, R->nbr_left, R->nbr_right, R->nbr_bottom,
                 R->nbr_top, R->cnr_tl, R->cnr_tr, R->cnr_bl, R->cnr_br};
    memcpy(out, v, sizeof v);
}
/* which: 0 xCenter 1 yCenter 2 Uc 3 Vc 4 rationalOmega 5 radius 6 wallTotalForceX 7 wallTotalForceY 8 totalTorque
 *        9 xCenterOld 10 yCenterOld 11 UcOld 12 VcOld 13 rationalOmegaOld */
double *p2_particle_ptr(p2_world *w, int which) {
    double *p[] = {w->xCenter, w->yCenter, w->Uc, w->Vc, w->rOmega, w->radius, w->wallTotalForceX, w->wallTotalForceY,
                   w->totalTorque, w->xCenterOld, w->yCenterOld, w->UcOld, w->VcOld, w->rOmegaOld};
    return p[which];
}
void p2_world_info(p2_world *w, int *dims, double *scal /* rhoAvg, errorU */, int *itc_err /* itc, error_flag */) {
    dims[0] = w->dims[0]; dims[1] = w->dims[1];
    scal[0] = w->rhoAvg; scal[1] = w->errorU;
    itc_err[0] = w->itc; itc_err[1] = w->error_flag;
}
void p2_get_params(p2_world *w, p2_params *p) { *p = w->p; }

static int inside(const p2_world *w, const p2_rank *R, int i, int j, const double *xc, const double *yc, int c) {
    /* ((i+i_start_global-xCenter)**2 + (j+j_start_global-yCenter)**2) .LE. radius**2 ; integer + integer - real */
    double dx = (double)(i + R->i_start) - xc[c], dy = (double)(j + R->j_start) - yc[c];
    return (dx * dx + dy * dy) <= w->radius[c] * w->radius[c];
}

/* initial() after the particle positions have been set by the caller, P4/initial.F90:79-199 */
void p2_initial(p2_world *w) {
    const int N = w->p.N;
    w->itc = 0; w->errorU = 100.0; w->error_flag = 0;
    for (int c = 0; c < N; ++c) {
        w->xCenterOld[c] = w->xCenter[c]; w->yCenterOld[c] = w->yCenter[c];
        w->Uc[c] = 0.0; w->Vc[c] = 0.0; w->rOmega[c] = 0.0; w->UcOld[c] = 0.0; w->VcOld[c] = 0.0; w->rOmegaOld[c] = 0.0;
    }
    w->omega[0] = 4.0 / 9.0;
    for (int a = 1; a <= 4; ++a) w->omega[a] = 1.0 / 9.0;
    for (int a = 5; a <= 8; ++a) w->omega[a] = 1.0 / 36.0;
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny;
        for (int j = 0; j <= ny + 1; ++j)
            for (int i = 0; i <= nx + 1; ++i) {
                OB(R, obst, i, j) = 0; OB(R, obstNew, i, j) = 0;
                for (int c = 0; c < N; ++c) if (inside(w, R, i, j, w->xCenter, w->yCenter, c)) OB(R, obst, i, j) = 1;
            }
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                S2(R, rho, i, j) = OB(R, obst, i, j) == 1 ? w->p.rhoSolid : w->p.rho0;
                S2(R, u, i, j) = 0.0; S2(R, v, i, j) = 0.0; S2(R, up, i, j) = 0.0; S2(R, vp, i, j) = 0.0;
            }
        memset(R->f, 0, Q9 * (size_t)(nx + 6) * (ny + 6) * 8);
        memset(R->f_post, 0, Q9 * (size_t)(nx + 4) * (ny + 4) * 8);
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i)
                if (OB(R, obst, i, j) == 0) {
                    double u = S2(R, u, i, j), v = S2(R, v, i, j), us2 = u * u + v * v;
                    for (int a = 0; a < Q9; ++a) {
                        double un = u * ex[a] + v * ey[a];
                        FI(R, a, i, j) = w->omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
                    }
                }
        /* ghost points, :150-184 (two layers of f and f_post; the third layer of f stays 0) */
        for (int j = -1; j <= ny + 2; ++j)
            for (int a = 0; a < Q9; ++a) {
                const double g = w->omega[a] * w->p.rho0;
                FI(R, a, 0, j) = g; FI(R, a, -1, j) = g; FI(R, a, nx + 1, j) = g; FI(R, a, nx + 2, j) = g;
                FP(R, a, 0, j) = g; FP(R, a, -1, j) = g; FP(R, a, nx + 1, j) = g; FP(R, a, nx + 2, j) = g;
            }
        for (int i = -1; i <= nx + 2; ++i)
            for (int a = 0; a < Q9; ++a) {
                const double g = w->omega[a] * w->p.rho0;
                FI(R, a, i, 0) = g; FI(R, a, i, -1) = g; FI(R, a, i, ny + 1) = g; FI(R, a, i, ny + 2) = g;
                FP(R, a, i, 0) = g; FP(R, a, i, -1) = g; FP(R, a, i, ny + 1) = g; FP(R, a, i, ny + 2) = g;
            }
    }
}

/* ---- collision(), P4/fluid.F90:1-72: one cell ---------------------------------------------------------- */
void p2_collide_cell(const double *f, double rho, double u, double v, double Snu, double Sq, double *fp) {
    double m[Q9], meq[Q9], s[Q9], mp[Q9];
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * f[5] + 2.0 * f[6] + 2.0 * f[7] + 2.0 * f[8];
    m[2] = 4.0 * f[0] - 2.0 * f[1] - 2.0 * f[2] - 2.0 * f[3] - 2.0 * f[4] + f[5] + f[6] + f[7] + f[8];
    m[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
    m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
    m[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
    m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
    m[7] = f[1] - f[2] + f[3] - f[4];
    m[8] = f[5] - f[6] + f[7] - f[8];
    meq[0] = rho;
    meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
    meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
    meq[3] = rho * u;
    meq[4] = -meq[3];
    meq[5] = rho * v;
    meq[6] = -meq[5];
    meq[7] = rho * (u * u - v * v);
    meq[8] = rho * (u * v);
    s[0] = 0.0; s[1] = Snu; s[2] = Snu; s[3] = 0.0; s[4] = Sq; s[5] = 0.0; s[6] = Sq; s[7] = Snu; s[8] = Snu;
    for (int a = 0; a < Q9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);
    fp[0] = (mp[0] - mp[1] + mp[2]) / 9.0;
    fp[1] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[3] / 6.0 - mp[4] / 6.0 + mp[7] * 0.25;
    fp[2] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[5] / 6.0 - mp[6] / 6.0 - mp[7] * 0.25;
    fp[3] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[3] / 6.0 + mp[4] / 6.0 + mp[7] * 0.25;
    fp[4] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[5] / 6.0 + mp[6] / 6.0 - mp[7] * 0.25;
    fp[5] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 + mp[8] * 0.25;
    fp[6] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 - mp[8] * 0.25;
    fp[7] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 + mp[8] * 0.25;
    fp[8] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 - mp[8] * 0.25;
}

void p2_collision(p2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 0) {
                    double fin[Q9], fout[Q9];
                    for (int a = 0; a < Q9; ++a) fin[a] = FI(R, a, i, j);
                    p2_collide_cell(fin, S2(R, rho, i, j), S2(R, u, i, j), S2(R, v, i, j), w->p.Snu, w->p.Sq, fout);
                    for (int a = 0; a < Q9; ++a) FP(R, a, i, j) = fout[a];
                }
    }
}

/* ---- send_all_fp() / send_all_f(), P4/message_send_all.F90: `depth`-deep halos, all 9 populations, faces then
 * corner squares.  which = 1: f_post (depth 2), which = 0: f (depth 3) ---------------------------------------- */
static double *cellp(p2_rank *R, int which, int i, int j) { return which ? &FP(R, 0, i, j) : &FI(R, 0, i, j); }
static void send_all(p2_world *w, int which) {
    const int depth = which ? 2 : 3;
    for (int r = 0; r < w->np; ++r) {
        p2_rank *S = &w->r[r];
        const int nx = S->nx, ny = S->ny;
        for (int l = 0; l < depth; ++l) {
            if (S->nbr_right >= 0) { p2_rank *D = &w->r[S->nbr_right];      /* column nx-l -> column -l */
                for (int j = 1; j <= ny; ++j) memcpy(cellp(D, which, -l, j), cellp(S, which, nx - l, j), Q9 * 8); }
            if (S->nbr_left >= 0) { p2_rank *D = &w->r[S->nbr_left];        /* column 1+l -> column nx+1+l */
                for (int j = 1; j <= ny; ++j) memcpy(cellp(D, which, D->nx + 1 + l, j), cellp(S, which, 1 + l, j), Q9 * 8); }
            if (S->nbr_top >= 0) { p2_rank *D = &w->r[S->nbr_top];          /* row ny-l -> row -l */
                for (int i = 1; i <= nx; ++i) memcpy(cellp(D, which, i, -l), cellp(S, which, i, ny - l), Q9 * 8); }
            if (S->nbr_bottom >= 0) { p2_rank *D = &w->r[S->nbr_bottom];    /* row 1+l -> row ny+1+l */
                for (int i = 1; i <= nx; ++i) memcpy(cellp(D, which, i, D->ny + 1 + l), cellp(S, which, i, 1 + l), Q9 * 8); }
        }
        for (int b = 0; b < depth; ++b)
            for (int a = 0; a < depth; ++a) {
                if (S->cnr_tr >= 0) { p2_rank *D = &w->r[S->cnr_tr];        /* (nx-d+1.., ny-d+1..) -> (-d+1.., -d+1..) */
                    memcpy(cellp(D, which, -depth + 1 + a, -depth + 1 + b), cellp(S, which, nx - depth + 1 + a, ny - depth + 1 + b), Q9 * 8); }
                if (S->cnr_tl >= 0) { p2_rank *D = &w->r[S->cnr_tl];        /* (1.., ny-d+1..) -> (nx+1.., -d+1..) */
                    memcpy(cellp(D, which, D->nx + 1 + a, -depth + 1 + b), cellp(S, which, 1 + a, ny - depth + 1 + b), Q9 * 8); }
                if (S->cnr_bl >= 0) { p2_rank *D = &w->r[S->cnr_bl];        /* (1.., 1..) -> (nx+1.., ny+1..) */
                    memcpy(cellp(D, which, D->nx + 1 + a, D->ny + 1 + b), cellp(S, which, 1 + a, 1 + b), Q9 * 8); }
                if (S->cnr_br >= 0) { p2_rank *D = &w->r[S->cnr_br];        /* (nx-d+1.., 1..) -> (-d+1.., ny+1..) */
                    memcpy(cellp(D, which, -depth + 1 + a, D->ny + 1 + b), cellp(S, which, nx - depth + 1 + a, 1 + b), Q9 * 8); }
            }
    }
}
void p2_send_all_fp(p2_world *w) { send_all(w, 1); }
void p2_send_all_f(p2_world *w) { send_all(w, 0); }

/* ---- streaming(), P4/fluid.F90:75-111: pull, skipped when the upstream node is solid --------------------- */
void p2_streaming(p2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                for (int a = 0; a < Q9; ++a) {
                    int ip = i - ex[a], jp = j - ey[a];
                    if (OB(R, obst, ip, jp) == 0) FI(R, a, i, j) = FP(R, a, ip, jp);
                }
    }
}

/* ---- bounceback(), P4/fluid.F90:115-161: the four channel walls ------------------------------------------ */
void p2_bounceback(p2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny;
        if (R->coords[0] == 0)
            for (int j = 1; j <= ny; ++j) { FI(R, 1, 1, j) = FP(R, 3, 1, j); FI(R, 5, 1, j) = FP(R, 7, 1, j); FI(R, 8, 1, j) = FP(R, 6, 1, j); }
        if (R->coords[0] == w->dims[0] - 1)
            for (int j = 1; j <= ny; ++j) { FI(R, 3, nx, j) = FP(R, 1, nx, j); FI(R, 6, nx, j) = FP(R, 8, nx, j); FI(R, 7, nx, j) = FP(R, 5, nx, j); }
        if (R->coords[1] == 0)
            for (int i = 1; i <= nx; ++i) { FI(R, 2, i, 1) = FP(R, 4, i, 1); FI(R, 5, i, 1) = FP(R, 7, i, 1); FI(R, 6, i, 1) = FP(R, 8, i, 1); }
        if (R->coords[1] == w->dims[1] - 1)
            for (int i = 1; i <= nx; ++i) { FI(R, 4, i, ny) = FP(R, 2, i, ny); FI(R, 7, i, ny) = FP(R, 5, i, ny); FI(R, 8, i, ny) = FP(R, 6, i, ny); }
        if (w->p.moving_walls) {
            /* moving top / bottom walls, P1/fluid.F90:127-145 (`movingFrame`; with U0 = 0 these are the `stationaryFrame` lines
             * :149-167 bit for bit, since -Uwall - 0 = -Uwall) */
            const double Uw = w->p.Uwall, U0 = w->p.Uframe;
            if (R->coords[1] == 0)
                for (int i = 1; i <= nx; ++i) {
                    FI(R, 5, i, 1) = FP(R, 7, i, 1) + (-Uw - U0) / 6.0;
                    FI(R, 6, i, 1) = FP(R, 8, i, 1) - (-Uw - U0) / 6.0;
                }
            if (R->coords[1] == w->dims[1] - 1)
                for (int i = 1; i <= nx; ++i) {
                    FI(R, 7, i, ny) = FP(R, 5, i, ny) - (Uw - U0) / 6.0;
                    FI(R, 8, i, ny) = FP(R, 6, i, ny) + (Uw - U0) / 6.0;
                }
        }
    }
}

/* ---- calQ, P4/particle_bounceback.F90:98-141: wall fraction along link alpha by bisection to 1e-9 (a
 * single-precision literal widened to fp64) ------------------------------------------------------------------ */
int p2_calQ(const p2_world *w, int c, double i, double j, int alpha, double *x0o, double *y0o, double *qo) {
    const double epsRadius = (double)1e-9f;
    const double xc = w->xCenter[c], yc = w->yCenter[c], rad = w->radius[c];
    double q = 0.5, qTemp = 0.5;
    double x0 = i + qTemp * (double)ex[alpha], y0 = j + qTemp * (double)ey[alpha];
    while (fabs(sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc)) - rad) >= epsRadius) {
        if (sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc)) > rad) {
            qTemp = qTemp / 2.0;
            x0 = x0 + qTemp * (double)ex[alpha]; y0 = y0 + qTemp * (double)ey[alpha];
            q = q + qTemp;
        } else if (sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc)) < rad) {
            qTemp = qTemp / 2.0;
            x0 = x0 - qTemp * (double)ex[alpha]; y0 = y0 - qTemp * (double)ey[alpha];
            q = q - qTemp;
        } else return -1;                       /* "error calQ!" (NaN) */
        if (qTemp == 0.0) return -1;            /* the reference would spin forever here */
    }
    *x0o = x0; *y0o = y0; *qo = q;
    return (q > 1.0 || q < 0.0) ? -2 : 0;       /* "error q!" */
}

/* rhoAvg = sum of rho over fluid nodes / number of fluid nodes, 2 Allreduce; mask = obst or obstNew */
static double fluid_average(p2_world *w, int use_new) {
    double total_rho = 0.0;
    long total_num = 0;
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        double s = 0.0;
        long n = 0;
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if ((use_new ? OB(R, obstNew, i, j) : OB(R, obst, i, j)) == 0) { s = s + S2(R, rho, i, j); n += 1; }
        total_rho += s; total_num += n;
    }
    return total_rho / (double)total_num;
}

/* one boundary link: fluid node (i,j) of rank R, direction alpha into particle c; P4/particle_bounceback.F90:58-75 */
void p2_bb_link(p2_world *w, p2_rank *R, int i, int j, int alpha, int c) {
    double x0, y0, q;
    int rc = p2_calQ(w, c, (double)(i + R->i_start), (double)(j + R->j_start), alpha, &x0, &y0, &q);
    if (rc) { w->error_flag = rc; return; }
    const double temp1 = -(y0 - w->yCenter[c]) * w->rOmega[c];
    const double temp2 = (x0 - w->xCenter[c]) * w->rOmega[c];
    const int ra = rr[alpha];
    const double Uc = w->Uc[c], Vc = w->Vc[c], rhoAvg = w->rhoAvg, om = w->omega[alpha];
    if (w->p.bb_linear) {                        /* P1/particle_bounceback.F90:66-76 */
        if (q < 0.5) {
            FI(R, ra, i, j) = 2.0 * q * FP(R, alpha, i, j)
                            + (1.0 - 2.0 * q) * FP(R, alpha, i - ex[alpha], j - ey[alpha])
                            + 6.0 * om * rhoAvg * (ex[ra] * (Uc + temp1) + ey[ra] * (Vc + temp2));
        } else if (q >= 0.5) {
            FI(R, ra, i, j) = 0.5 / q * FP(R, alpha, i, j)
                            + (1.0 - 0.50 / q) * FP(R, ra, i, j)
                            + 3.0 * om * rhoAvg / q * (ex[ra] * (Uc + temp1) + ey[ra] * (Vc + temp2));
        }
        return;
    }
    if (q < 0.5) {
        FI(R, ra, i, j) = q * (1.0 + 2.0 * q) * FP(R, alpha, i, j)
                        + (1.0 - 4.0 * q * q) * FP(R, alpha, i - ex[alpha], j - ey[alpha])
                        - q * (1.0 - 2.0 * q) * FP(R, alpha, i - 2 * ex[alpha], j - 2 * ey[alpha])
                        + 6.0 * om * rhoAvg * (ex[ra] * (Uc + temp1) + ey[ra] * (Vc + temp2));
    } else if (q >= 0.5) {
        FI(R, ra, i, j) = FP(R, alpha, i, j) / q / (1.0 + 2.0 * q)
                        + FP(R, ra, i, j) * (2.0 * q - 1.0) / q
                        - FP(R, ra, i - ex[alpha], j - ey[alpha]) * (2.0 * q - 1.0) / (2.0 * q + 1.0)
                        + 6.0 * om * rhoAvg / q / (1.0 + 2.0 * q) * (ex[ra] * (Uc + temp1) + ey[ra] * (Vc + temp2));
    }
}

/* ---- bounceback_particle(), P4/particle_bounceback.F90:1-96 ------------------------------------------------- */
void p2_bounceback_particle(p2_world *w) {
    w->rhoAvg = fluid_average(w, 0);
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 0)
                    for (int a = 0; a < Q9; ++a) {
                        int ip = i + ex[a], jp = j + ey[a];
                        if (OB(R, obst, ip, jp) == 1) {
                            int myFlag = 0;
                            for (int c = 0; c < w->p.N; ++c)
                                if (inside(w, R, ip, jp, w->xCenter, w->yCenter, c)) { myFlag = 1; p2_bb_link(w, R, i, j, a, c); }
                            if (!myFlag) w->error_flag = -3;       /* "Did not find the center owning the boundary points!" */
                        }
                    }
    }
}

/* ---- macro(), P4/fluid.F90:164-184 ------------------------------------------------------------------------------ */
void p2_macro_cell(const double *f, double *out) {
    out[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    out[1] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / out[0];
    out[2] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / out[0];
}
void p2_macro(p2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 0) {
                    double f[Q9], o[3];
                    for (int a = 0; a < Q9; ++a) f[a] = FI(R, a, i, j);
                    p2_macro_cell(f, o);
                    S2(R, rho, i, j) = o[0]; S2(R, u, i, j) = o[1]; S2(R, v, i, j) = o[2];
                }
    }
}

/* momentum exchange of one link, P4/particle_force.F90:51-62 (Wen et al., JCP 2014) */
int p2_force_link(p2_world *w, p2_rank *R, int i, int j, int alpha, int c, double *out /* Fx, Fy, torque */) {
    double x0, y0, q;
    int rc = p2_calQ(w, c, (double)(i + R->i_start), (double)(j + R->j_start), alpha, &x0, &y0, &q);
    if (rc) return rc;
    const double temp1 = -(y0 - w->yCenter[c]) * w->rOmega[c];
    const double temp2 = (x0 - w->xCenter[c]) * w->rOmega[c];
    const int ra = rr[alpha];
    out[0] = (ex[alpha] - w->Uc[c] - temp1) * FP(R, alpha, i, j) - (ex[ra] - w->Uc[c] - temp1) * FI(R, ra, i, j);
    out[1] = (ey[alpha] - w->Vc[c] - temp2) * FP(R, alpha, i, j) - (ey[ra] - w->Vc[c] - temp2) * FI(R, ra, i, j);
    out[2] = (x0 - w->xCenter[c]) * out[1] - (y0 - w->yCenter[c]) * out[0];
    return 0;
}

/* the owner's additions to one particle's force: spring repulsion particle-particle and particle-wall,
 * buoyancy-corrected weight; P4/particle_force.F90:95-190 */
int p2_particle_forces(p2_world *w, int c, double *Fx, double *Fy) {
    const p2_params *p = &w->p;
    const int N = p->N;
    const double rad = w->radius[c];
    double forceScale = p->Pi * (rad * rad) * (p->rhoSolid - w->rhoAvg) * p->gravity / p->stiffParticle;
    double Fxij = 0.0, Fyij = 0.0;
    for (int c2 = 0; c2 < N; ++c2)
        if (c2 != c) {
            const double ddx = w->xCenter[c] - w->xCenter[c2], ddy = w->yCenter[c] - w->yCenter[c2];
            const double dij = sqrt(ddx * ddx + ddy * ddy);
            if (dij >= (rad + w->radius[c2] + p->thresholdParticle)) {
            } else if (dij < (rad + w->radius[c2] + p->thresholdParticle) && dij >= (rad + w->radius[c2])) {
                const double t = (dij - rad - w->radius[c2] - p->thresholdParticle) / p->thresholdParticle;
                Fxij = Fxij + forceScale * (t * t) * (w->xCenter[c] - w->xCenter[c2]) / dij;
                Fyij = Fyij + forceScale * (t * t) * (w->yCenter[c] - w->yCenter[c2]) / dij;
            } else return -4;                    /* 'Particle-particle interpenetration!' -> MPI_Abort */
        }
    double Fwx = 0.0, Fwy = 0.0;
    forceScale = p->Pi * (rad * rad) * (p->rhoSolid - p->rho0) * p->gravity / p->stiffWall;
    double dw = w->yCenter[c] - rad - 1.0;                                   /* bottom wall */
    if (dw < 0) return -5;
    else if (dw < p->thresholdWall) { const double t = (dw - p->thresholdWall) / p->thresholdWall; Fwy = Fwy + forceScale * (t * t); }
    dw = w->xCenter[c] - rad - 1.0;                                          /* left wall */
    if (dw < 0) return -5;
    else if (dw < p->thresholdWall) { const double t = (dw - p->thresholdWall) / p->thresholdWall; Fwx = Fwx + forceScale * (t * t); }
    dw = (double)p->total_nx - w->xCenter[c] - rad;                          /* right wall */
    if (dw < 0) return -5;
    else if (dw < p->thresholdWall) { const double t = (dw - p->thresholdWall) / p->thresholdWall; Fwx = Fwx - forceScale * (t * t); }
    *Fx = w->wallTotalForceX[c] + Fxij + Fwx;                                /* :184-190 */
    *Fy = w->wallTotalForceY[c] - (p->rhoSolid - w->rhoAvg) * p->Pi * (p->radius0 * p->radius0) * p->gravity + Fyij + Fwy;
    return 0;
}

/* ---- calForce(), P4/particle_force.F90:1-212 ----------------------------------------------------------------- */
void p2_calForce(p2_world *w) {
    const int N = w->p.N;
    for (int c = 0; c < N; ++c) { w->wallTotalForceX[c] = 0.0; w->wallTotalForceY[c] = 0.0; w->totalTorque[c] = 0.0; }
    double *fx = calloc((size_t)N, 8), *fy = calloc((size_t)N, 8), *tq = calloc((size_t)N, 8);
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        memset(fx, 0, (size_t)N * 8); memset(fy, 0, (size_t)N * 8); memset(tq, 0, (size_t)N * 8);
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 0)
                    for (int a = 1; a < Q9; ++a) {
                        int ip = i + ex[a], jp = j + ey[a];
                        if (OB(R, obst, ip, jp) == 1) {
                            int myFlag = 0;
                            for (int c = 0; c < N; ++c)
                                if (inside(w, R, ip, jp, w->xCenter, w->yCenter, c)) {
                                    myFlag = 1;
                                    double o[3];
                                    int rc = p2_force_link(w, R, i, j, a, c, o);
                                    if (rc) { w->error_flag = rc; continue; }
                                    fx[c] = fx[c] + o[0]; fy[c] = fy[c] + o[1]; tq[c] = tq[c] + o[2];
                                }
                            if (!myFlag) w->error_flag = -3;
                        }
                    }
        for (int c = 0; c < N; ++c) {           /* MPI_Allreduce(SUM) x3 as a rank-ordered sum, :90-92 */
            w->wallTotalForceX[c] += fx[c]; w->wallTotalForceY[c] += fy[c]; w->totalTorque[c] += tq[c];
        }
    }
    free(fx); free(fy); free(tq);
    for (int c = 0; c < N; ++c) {               /* every particle has exactly one owner (local_mask) */
        double Fx, Fy;
        int rc = p2_particle_forces(w, c, &Fx, &Fy);
        if (rc) { w->error_flag = rc; continue; }
        w->wallTotalForceX[c] = Fx; w->wallTotalForceY[c] = Fy;
    }
}

/* explicit kinematics of one particle, P4/particle_update.F90:37-50; inertia uses radius**4.0d0 (libm pow) */
void p2_particle_advance(const p2_params *p, double Fx, double Fy, double torque, double radius, double xOld, double yOld,
                         double UOld, double VOld, double omOld, double *out /* x, y, U, V, omega */) {
    const double ax = Fx / p->Pi / (p->radius0 * p->radius0) / p->rhoSolid;
    const double ay = Fy / p->Pi / (p->radius0 * p->radius0) / p->rhoSolid;
    const double aOmega = torque / (0.5 * p->rhoSolid * p->Pi * pow(radius, 4.0));
    out[2] = UOld + ax;
    out[3] = VOld + ay;
    out[4] = omOld + aOmega;
    out[0] = xOld + UOld + 0.5 * ax;
    out[1] = yOld + VOld + 0.5 * ay;
}

/* refill of one newly uncovered node from particle c, P4/particle_update.F90:137-191 */
int p2_refill_cell(p2_world *w, p2_rank *R, int i, int j, int c) {
    double outNormal = 0.0;
    int ec = 0;
    const double dx = (double)(i + R->i_start) - w->xCenter[c], dy = (double)(j + R->j_start) - w->yCenter[c];
    for (int a = 1; a < Q9; ++a) {
        double tempNormal = (dx * ex[a] + dy * ey[a]) / sqrt(dx * dx + dy * dy);
        if (tempNormal > outNormal) { outNormal = tempNormal; ec = a; }
    }
    if (ec == 0) return -6;
    for (int a = 0; a < Q9; ++a)
        FI(R, a, i, j) = 3.0 * FI(R, a, i + ex[ec], j + ey[ec]) - 3.0 * FI(R, a, i + 2 * ex[ec], j + 2 * ey[ec])
                       + FI(R, a, i + 3 * ex[ec], j + 3 * ey[ec]);
    double f[Q9], m[Q9];
    for (int a = 0; a < Q9; ++a) f[a] = FI(R, a, i, j);
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * f[5] + 2.0 * f[6] + 2.0 * f[7] + 2.0 * f[8];
    m[2] = 4.0 * f[0] - 2.0 * f[1] - 2.0 * f[2] - 2.0 * f[3] - 2.0 * f[4] + f[5] + f[6] + f[7] + f[8];
    m[3] = w->rhoAvg * (w->Uc[c] - ((double)(j + R->j_start) - w->yCenter[c]) * w->rOmega[c]);
    m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
    m[5] = w->rhoAvg * (w->Vc[c] + ((double)(i + R->i_start) - w->xCenter[c]) * w->rOmega[c]);
    m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
    m[7] = f[1] - f[2] + f[3] - f[4];
    m[8] = f[5] - f[6] + f[7] - f[8];
    f[0] = (m[0] - m[1] + m[2]) / 9.0;
    f[1] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 + m[3] / 6.0 - m[4] / 6.0 + m[7] * 0.25;
    f[2] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 + m[5] / 6.0 - m[6] / 6.0 - m[7] * 0.25;
    f[3] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 - m[3] / 6.0 + m[4] / 6.0 + m[7] * 0.25;
    f[4] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 - m[5] / 6.0 + m[6] / 6.0 - m[7] * 0.25;
    f[5] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 + m[3] / 6.0 + m[4] / 12.0 + m[5] / 6.0 + m[6] / 12.0 + m[8] * 0.25;
    f[6] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 - m[3] / 6.0 - m[4] / 12.0 + m[5] / 6.0 + m[6] / 12.0 - m[8] * 0.25;
    f[7] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 - m[3] / 6.0 - m[4] / 12.0 - m[5] / 6.0 - m[6] / 12.0 + m[8] * 0.25;
    f[8] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 + m[3] / 6.0 + m[4] / 12.0 - m[5] / 6.0 - m[6] / 12.0 - m[8] * 0.25;
    for (int a = 0; a < Q9; ++a) FI(R, a, i, j) = f[a];
    double o[3];
    p2_macro_cell(f, o);
    S2(R, rho, i, j) = o[0]; S2(R, u, i, j) = o[1]; S2(R, v, i, j) = o[2];
    return 0;
}

/* ---- updateCenter(), P4/particle_update.F90:1-209 (+ update_particle_mask, message_particle.F90:470-497) ---- */
void p2_updateCenter(p2_world *w) {
    const int N = w->p.N;
    for (int c = 0; c < N; ++c) {
        w->xCenterOld[c] = w->xCenter[c]; w->yCenterOld[c] = w->yCenter[c];
        w->UcOld[c] = w->Uc[c]; w->VcOld[c] = w->Vc[c]; w->rOmegaOld[c] = w->rOmega[c];
    }
    for (int c = 0; c < N; ++c) {
        double o[5];
        p2_particle_advance(&w->p, w->wallTotalForceX[c], w->wallTotalForceY[c], w->totalTorque[c], w->radius[c], w->xCenterOld[c],
                            w->yCenterOld[c], w->UcOld[c], w->VcOld[c], w->rOmegaOld[c], o);
        w->xCenter[c] = o[0]; w->yCenter[c] = o[1]; w->Uc[c] = o[2]; w->Vc[c] = o[3]; w->rOmega[c] = o[4];
    }
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        for (int j = 0; j <= R->ny + 1; ++j)
            for (int i = 0; i <= R->nx + 1; ++i) {
                OB(R, obstNew, i, j) = 0;
                for (int c = 0; c < N; ++c) if (inside(w, R, i, j, w->xCenter, w->yCenter, c)) OB(R, obstNew, i, j) = 1;
            }
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obstNew, i, j) == 1) { S2(R, rho, i, j) = w->p.rhoSolid; S2(R, u, i, j) = 0.0; S2(R, v, i, j) = 0.0; }
    }
    w->rhoAvg = fluid_average(w, 1);             /* over obstNew; freshly uncovered nodes still carry rhoSolid */
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 1 && OB(R, obstNew, i, j) == 0) {
                    int myFlag = 0;
                    for (int c = 0; c < N; ++c)
                        if (inside(w, R, i, j, w->xCenterOld, w->yCenterOld, c)) {
                            myFlag = 1;
                            int rc = p2_refill_cell(w, R, i, j, c);
                            if (rc) w->error_flag = rc;
                        }
                    if (!myFlag) w->error_flag = -7;
                }
        memcpy(R->obst, R->obstNew, (size_t)(R->nx + 2) * (R->ny + 2) * sizeof(int));
    }
}

/* ---- check(), P4/fluid.F90:187-221 ------------------------------------------------------------------------------ */
double p2_check(p2_world *w) {
    double t1 = 0.0, t2 = 0.0;
    for (int r = 0; r < w->np; ++r) {
        p2_rank *R = &w->r[r];
        double e1 = 0.0, e2 = 0.0;
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                if (OB(R, obst, i, j) == 0) {
                    double u = S2(R, u, i, j), v = S2(R, v, i, j);
                    e1 = e1 + (u - S2(R, up, i, j)) * (u - S2(R, up, i, j)) + (v - S2(R, vp, i, j)) * (v - S2(R, vp, i, j));
                    e2 = e2 + u * u + v * v;
                    S2(R, up, i, j) = u; S2(R, vp, i, j) = v;
                }
        t1 += e1; t2 += e2;
    }
    w->errorU = sqrt(t1) / sqrt(t2);
    return w->errorU;
}

/* n iterations of the driver loop body, P4/main.F90:35-73 (check/output every 500 left to the caller) */
void p2_step(p2_world *w, int n) {
    for (int s = 0; s < n; ++s) {
        p2_collision(w);
        p2_send_all_fp(w);
        p2_streaming(w);
        p2_bounceback(w);
        p2_bounceback_particle(w);
        p2_macro(w);
        p2_calForce(w);
        p2_send_all_f(w);
        w->itc += 1;
        p2_updateCenter(w);
    }
}

/* ---- small entry points for the per-link / per-node golden-vector tests ---------------------------------- */
void p2_set_rhoAvg(p2_world *w, double v) { w->rhoAvg = v; }
void p2_bb_link_r(p2_world *w, int r, int i, int j, int alpha, int c) { p2_bb_link(w, &w->r[r], i, j, alpha, c); }
int p2_force_link_r(p2_world *w, int r, int i, int j, int alpha, int c, double *out) { return p2_force_link(w, &w->r[r], i, j, alpha, c, out); }
int p2_refill_cell_r(p2_world *w, int r, int i, int j, int c) { return p2_refill_cell(w, &w->r[r], i, j, c); }
