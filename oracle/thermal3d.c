/*
 * oracle/thermal3d.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's thermal
 * double-distribution hot path (cheryli/MGLC, MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90,
 * "B3" below; its split copy mpi_blocked/evolution_{f,g}.F90 carries the same bodies): D3Q19 MRT flow with
 * Boussinesq + Coriolis forcing, D3Q7 MRT temperature, side-heated cavity.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load this.
 *
 * PARITY PIN: the Fortran+MPI program cannot be built in this image.  The per-cell arithmetic restated
 * here (collision + forcing, macro, collisionT, both equilibria, the module parameters) is checked bit for
 * bit against vectors obtained by machine-evaluating the reference's own source text
 * (tests/golden/make_golden_fortran.py -> tests/golden/ref_fortran_kernels.npz); the copy-type subroutines are
 * pinned the same way as whole arrays (make_golden_thermal3d_fields.py -> ref_fortran_thermal3d_fields.npz:
 * streamingT B3:1081-1094, bounceback B3:900-980, bouncebackT B3:1106-1207 with the benchmark-cavity and the
 * RB-convection macro sets for 13 block positions, the check() sums B3:1242-1264); the exchange is MPI calls and is
 * checked by construction tests (tests/test_oracle_thermal.py).  Whole run: the SEQUENTIAL program 3d/seq/bouyancy3d.F90 with its
 * shipped macro set (and again with its RB-convection set) is evaluated from its text on 6 x 5 x 4 -- parameters, initial() and the driver loop for 1, 2, 10, 12
 * iterations with check() (make_golden_thermal3d_seq_run.py -> ref_fortran_thermal3d_seq_run.npz) -- and this file reproduces its
 * f, g, rho, u, v, w, T, Fx, Fy, Fz bit for bit on 1..8 emulated ranks.
 *
 * Layout is the reference's (B3:483-492): f(0:18,nx,ny,nz), f_post(0:18,0:nx+1,0:ny+1,0:nz+1),
 * g(0:6,nx,ny,nz), g_post(0:6,0:nx+1,...), rho,u,v,w,T,Fx,Fy,Fz,up,vp,wp,Tp (nx,ny,nz), column-major.
 * Left-to-right expressions, divisions stay divisions, -ffp-contract=off.  One process emulates all ranks.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q 19
#define QT 7

void orc_dims_create(int np, int dims[3]);                                        /* lid3d.c */
void orc_decompose_1d(int total_n, int rank, int np, int *local_n, int *start);   /* lid3d.c */

static const int ex[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int ey[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int ez[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};

enum { TH_ADIABATIC = 0, TH_CONST_HOT = 1, TH_CONST_COLD = 2 };

typedef struct th_params {
    double Rayleigh, Prandtl, Mach, Ekman, Thot, Tcold, Tref;
    double tauf, viscosity, diffusivity, omegaRatating, paraA, gBeta1, gBeta, Snu, Sq, Qd, Qnu;
} th_params;

typedef struct th_rank {
    int nx, ny, nz, coords[3], start[3];
    int nbr_surface[7], nbr_line[Q];
    double *f, *f_post, *g, *g_post;
    double *rho, *u, *v, *w, *T, *Fx, *Fy, *Fz, *up, *vp, *wp, *Tp;
} th_rank;

typedef struct th_world {
    int total[3], dims[3], np, itc;
    int bcT[6];           /* +x,-x,+y,-y,+z,-z : TH_* */
    th_params p;
    double errorU, errorT;
    th_rank *r;
} th_world;

#define F(R, a, i, j, k) ((R)->f[(a) + Q * ((size_t)((i)-1) + (size_t)(R)->nx * ((size_t)((j)-1) + (size_t)(R)->ny * (size_t)((k)-1)))])
#define FP(R, a, i, j, k) ((R)->f_post[(a) + Q * ((size_t)(i) + (size_t)((R)->nx + 2) * ((size_t)(j) + (size_t)((R)->ny + 2) * (size_t)(k)))])
#define G(R, a, i, j, k) ((R)->g[(a) + QT * ((size_t)((i)-1) + (size_t)(R)->nx * ((size_t)((j)-1) + (size_t)(R)->ny * (size_t)((k)-1)))])
#define GP(R, a, i, j, k) ((R)->g_post[(a) + QT * ((size_t)(i) + (size_t)((R)->nx + 2) * ((size_t)(j) + (size_t)((R)->ny + 2) * (size_t)(k)))])
#define S(R, A, i, j, k) ((R)->A[(size_t)((i)-1) + (size_t)(R)->nx * ((size_t)((j)-1) + (size_t)(R)->ny * (size_t)((k)-1))])

/* module commondata parameters, B3:30-47,73-74 */
void th_make_params(int total_nz, double Rayleigh, double Prandtl, double Mach, double Ekman, double Thot,
                    double Tcold, double Tref, th_params *p) {
    p->Rayleigh = Rayleigh; p->Prandtl = Prandtl; p->Mach = Mach; p->Ekman = Ekman;
    p->Thot = Thot; p->Tcold = Tcold; p->Tref = Tref;
    p->tauf = 0.5 + Mach * (double)total_nz * sqrt(3.0 * Prandtl / Rayleigh);
    p->viscosity = (p->tauf - 0.5) / 3.0;
    p->diffusivity = p->viscosity / Prandtl;
    p->omegaRatating = p->viscosity / 2.0 / Ekman / (double)(total_nz * total_nz);
    p->paraA = 42.0 * sqrt(3.0) * p->diffusivity - 6.0;
    p->gBeta1 = Rayleigh * p->viscosity * p->diffusivity / (double)total_nz;
    p->gBeta = p->gBeta1 / (double)(total_nz * total_nz);
    p->Snu = 1.0 / p->tauf;
    p->Sq = 8.0 * (2.0 * p->tauf - 1.0) / (8.0 * p->tauf - 1.0);
    p->Qd = 3.0 - sqrt(3.0);
    p->Qnu = 4.0 * sqrt(3.0) - 6.0;
}

static int cart_rank(const int dims[3], const int c[3]) {
    for (int d = 0; d < 3; ++d) if (c[d] < 0 || c[d] >= dims[d]) return -1;
    return (c[0] * dims[1] + c[1]) * dims[2] + c[2];
}

/* bc_or_null: thermal wall kinds per face; NULL = the shipped benchmarkCavity set (B3:15-19):
 * y walls constant T (hot at j = 1, cold at j = ny), x and z walls adiabatic */
th_world *th_world_create(int tnx, int tny, int tnz, int np, const int *dims_or_null, const int *bc_or_null,
                          double Rayleigh, double Prandtl, double Mach, double Ekman) {
    th_world *w = (th_world *)calloc(1, sizeof *w);
    w->total[0] = tnx; w->total[1] = tny; w->total[2] = tnz; w->np = np;
    if (dims_or_null && dims_or_null[0] > 0) memcpy(w->dims, dims_or_null, sizeof w->dims);
    else orc_dims_create(np, w->dims);
    static const int shipped[6] = {TH_ADIABATIC, TH_ADIABATIC, TH_CONST_COLD, TH_CONST_HOT, TH_ADIABATIC, TH_ADIABATIC};
    memcpy(w->bcT, bc_or_null ? bc_or_null : shipped, sizeof w->bcT);
    th_make_params(tnz, Rayleigh, Prandtl, Mach, Ekman, 1.0, 0.0, 0.0, &w->p);
    w->r = (th_rank *)calloc((size_t)np, sizeof(th_rank));
    for (int c0 = 0; c0 < w->dims[0]; ++c0)
    for (int c1 = 0; c1 < w->dims[1]; ++c1)
    for (int c2 = 0; c2 < w->dims[2]; ++c2) {
        int c[3] = {c0, c1, c2};
        th_rank *R = &w->r[cart_rank(w->dims, c)];
        memcpy(R->coords, c, sizeof c);
        orc_decompose_1d(tnx, c0, w->dims[0], &R->nx, &R->start[0]);
        orc_decompose_1d(tny, c1, w->dims[1], &R->ny, &R->start[1]);
        orc_decompose_1d(tnz, c2, w->dims[2], &R->nz, &R->start[2]);
        for (int d = 0; d < 3; ++d) {                 /* B3:153-155 */
            int p[3] = {c0, c1, c2}, m[3] = {c0, c1, c2};
            p[d] += 1; m[d] -= 1;
            R->nbr_surface[2 * d + 1] = cart_rank(w->dims, p);
            R->nbr_surface[2 * d + 2] = cart_rank(w->dims, m);
        }
        for (int a = 7; a < Q; ++a) {                 /* B3:352-404 */
            int n[3] = {c0 + ex[a], c1 + ey[a], c2 + ez[a]};
            R->nbr_line[a] = cart_rank(w->dims, n);
        }
        size_t n = (size_t)R->nx * R->ny * R->nz, nh = (size_t)(R->nx + 2) * (R->ny + 2) * (R->nz + 2);
        R->f = calloc(Q * n, 8); R->f_post = calloc(Q * nh, 8);
        R->g = calloc(QT * n, 8); R->g_post = calloc(QT * nh, 8);
        double **fields[] = {&R->rho, &R->u, &R->v, &R->w, &R->T, &R->Fx, &R->Fy, &R->Fz, &R->up, &R->vp, &R->wp, &R->Tp};
        for (size_t q = 0; q < sizeof fields / sizeof *fields; ++q) *fields[q] = calloc(n, 8);
    }
    return w;
}

void th_world_destroy(th_world *w) {
    if (!w) return;
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        double *p[] = {R->f, R->f_post, R->g, R->g_post, R->rho, R->u, R->v, R->w, R->T, R->Fx, R->Fy, R->Fz, R->up, R->vp, R->wp, R->Tp};
        for (size_t q = 0; q < sizeof p / sizeof *p; ++q) free(p[q]);
    }
    free(w->r); free(w);
}

/* which: 0 f, 1 f_post, 2 g, 3 g_post, 4 rho, 5 u, 6 v, 7 w, 8 T, 9 Fx, 10 Fy, 11 Fz, 12 up, 13 vp, 14 wp, 15 Tp */
double *th_rank_ptr(th_world *w, int r, int which) {
    th_rank *R = &w->r[r];
    double *p[] = {R->f, R->f_post, R->g, R->g_post, R->rho, R->u, R->v, R->w, R->T, R->Fx, R->Fy, R->Fz, R->up, R->vp, R->wp, R->Tp};
    return (which >= 0 && which < 16) ? p[which] : NULL;
}
void th_rank_info(th_world *w, int r, int *out /*[27]*/) {
    th_rank *R = &w->r[r];
    out[0] = R->nx; out[1] = R->ny; out[2] = R->nz;
    for (int d = 0; d < 3; ++d) { out[3 + d] = R->coords[d]; out[6 + d] = R->start[d]; }
    for (int s = 1; s <= 6; ++s) out[8 + s] = R->nbr_surface[s];
    for (int a = 7; a < Q; ++a) out[8 + a] = R->nbr_line[a];
}
void th_world_info(th_world *w, int *dims, th_params *p, int *bcT) {
    memcpy(dims, w->dims, sizeof w->dims);
    *p = w->p;
    memcpy(bcT, w->bcT, sizeof w->bcT);
}

/* ---- equilibria of initial(), B3:496-507, 589-597 ---------------------------------------------- */
void th_feq_cell(double rho, double u, double v, double w, double *f) {
    double omega[Q];
    omega[0] = 1.0 / 3.0;
    for (int a = 1; a <= 6; ++a) omega[a] = 1.0 / 18.0;
    for (int a = 7; a <= 18; ++a) omega[a] = 1.0 / 36.0;
    double us2 = u * u + v * v + w * w;
    for (int a = 0; a < Q; ++a) {
        double un = u * ex[a] + v * ey[a] + w * ez[a];
        f[a] = rho * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
    }
}
void th_geq_cell(double T, double u, double v, double w, double paraA, double *g) {
    double omegaT[QT];
    omegaT[0] = (1.0 - paraA) / 7.0;
    for (int a = 1; a <= 6; ++a) omegaT[a] = (paraA + 6.0) / 42.0;
    for (int a = 0; a < QT; ++a) {
        double unT = u * ex[a] + v * ey[a] + w * ez[a];
        g[a] = omegaT[a] * T * (1.0 + 21.0 / (6.0 + paraA) * unT);
    }
}

/* ---- initial(), B3:409-638 ------------------------------------------------------------------------ */
void th_initial(th_world *w) {
    w->itc = 0; w->errorU = 100.0; w->errorT = 100.0;
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny, nz = R->nz;
        size_t n = (size_t)nx * ny * nz, nh = (size_t)(nx + 2) * (ny + 2) * (nz + 2);
        for (size_t q = 0; q < n; ++q) {
            R->rho[q] = 1.0; R->u[q] = 0.0; R->v[q] = 0.0; R->w[q] = 0.0; R->T[q] = 0.0;
            R->up[q] = 0.0; R->vp[q] = 0.0; R->wp[q] = 0.0; R->Tp[q] = 0.0;
        }
        /* wall-adjacent layers of constant-temperature walls start at the wall temperature, :542-587
         * (LeftRightWallsConstT first, then TopBottomPlatesConstT; x walls have no constant-T option) */
        for (int pass = 0; pass < 2; ++pass) {
            const int axis = pass == 0 ? 1 : 2;
            for (int side = 1; side >= 0; --side) {          /* minus face first (hot wall in the shipped set) */
                const int face = 2 * axis + side;
                if (w->bcT[face] == TH_ADIABATIC) continue;
                const int at_wall = side ? (R->coords[axis] == 0) : (R->coords[axis] == w->dims[axis] - 1);
                if (!at_wall) continue;
                const double Tw = w->bcT[face] == TH_CONST_HOT ? w->p.Thot : w->p.Tcold;
                if (axis == 1) { const int j = side ? 1 : ny; for (int k = 1; k <= nz; ++k) for (int i = 1; i <= nx; ++i) S(R, T, i, j, k) = Tw; }
                else { const int k = side ? 1 : nz; for (int j = 1; j <= ny; ++j) for (int i = 1; i <= nx; ++i) S(R, T, i, j, k) = Tw; }
            }
        }
        for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            th_feq_cell(S(R, rho, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k), S(R, w, i, j, k), &F(R, 0, i, j, k));
            th_geq_cell(S(R, T, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k), S(R, w, i, j, k), w->p.paraA, &G(R, 0, i, j, k));
        }
        memset(R->f_post, 0, Q * nh * 8);        /* f_post = 0, g_post = 0, :624-625 */
        memset(R->g_post, 0, QT * nh * 8);
    }
}

/* ---- collision(), B3:640-864: one cell ------------------------------------------------------------- */
void th_collide_cell(const double *f, double rho, double u, double v, double w, double T, const th_params *p,
                     double *fp, double *Fxyz) {
    double m[Q], meq[Q], s[Q], fs[Q], mp[Q];
    /* forward transform, :656-700 (note the groupings differ from the lid driver's) */
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6]
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[1] = -30.0 * f[0] - 11.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + 8.0 * (f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18]);
    m[2] = 12.0 * f[0] - 4.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[3] = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[4] = -4.0 * (f[1] - f[2]) + f[7] - f[8] + f[9] - f[10]
         + f[11] - f[12] + f[13] - f[14];
    m[5] = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[6] = -4.0 * (f[3] - f[4]) + f[7] + f[8] - f[9] - f[10]
         + f[15] - f[16] + f[17] - f[18];
    m[7] = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[8] = -4.0 * (f[5] - f[6]) + f[11] + f[12] - f[13] - f[14]
         + f[15] + f[16] - f[17] - f[18];
    m[9] = 2.0 * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6]
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[10] = -4.0 * (f[1] + f[2]) + 2.0 * (f[3] + f[4] + f[5] + f[6])
          + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[11] = f[3] + f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] - (f[11] + f[12] + f[13] + f[14]);
    m[12] = -2.0 * (f[3] + f[4] - f[5] - f[6]) + (f[7] + f[8] + f[9] + f[10]) - (f[11] + f[12] + f[13] + f[14]);
    m[13] = f[7] - f[8] - f[9] + f[10];
    m[14] = f[15] - f[16] - f[17] + f[18];
    m[15] = f[11] - f[12] - f[13] + f[14];
    m[16] = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    m[17] = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    m[18] = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];
    /* equilibrium moments, :703-721 (meq(12) = -0.5*meq(11): WITH rho, unlike the lid driver) */
    meq[0] = rho;
    meq[1] = -11.0 * rho + 19.0 * rho * (u * u + v * v + w * w);
    meq[2] = 3.0 * rho - 11.0 / 2.0 * rho * (u * u + v * v + w * w);
    meq[3] = rho * u;
    meq[4] = -2.0 / 3.0 * meq[3];
    meq[5] = rho * v;
    meq[6] = -2.0 / 3.0 * meq[5];
    meq[7] = rho * w;
    meq[8] = -2.0 / 3.0 * meq[7];
    meq[9] = rho * (2.0 * u * u - v * v - w * w);
    meq[10] = -0.5 * meq[9];
    meq[11] = rho * (v * v - w * w);
    meq[12] = -0.5 * meq[11];
    meq[13] = rho * (u * v);
    meq[14] = rho * (v * w);
    meq[15] = rho * (w * u);
    meq[16] = 0.0; meq[17] = 0.0; meq[18] = 0.0;
    /* relaxation rates, :723-741 */
    const double Snu = p->Snu, Sq = p->Sq;
    s[0] = 0.0; s[1] = Snu; s[2] = Snu; s[3] = 0.0; s[4] = Sq; s[5] = 0.0; s[6] = Sq; s[7] = 0.0; s[8] = Sq;
    s[9] = Snu; s[10] = Snu; s[11] = Snu; s[12] = Snu; s[13] = Snu; s[14] = Snu; s[15] = Snu;
    s[16] = Sq; s[17] = Sq; s[18] = Sq;
    /* body force from the PREVIOUS step's fields: Coriolis + Boussinesq, :743-745 */
    const double Fx = -2.0 * rho * v * p->omegaRatating;
    const double Fy = 2.0 * rho * u * p->omegaRatating;
    const double Fz = rho * p->gBeta * (T - p->Tref);
    Fxyz[0] = Fx; Fxyz[1] = Fy; Fxyz[2] = Fz;
    /* moment-space source, :747-765 */
    fs[0] = 0.0;
    fs[1] = 38.0 * (u * Fx + v * Fy + w * Fz);
    fs[2] = -11.0 * (u * Fx + v * Fy + w * Fz);
    fs[3] = Fx;
    fs[4] = -2.0 / 3.0 * Fx;
    fs[5] = Fy;
    fs[6] = -2.0 / 3.0 * Fy;
    fs[7] = Fz;
    fs[8] = -2.0 / 3.0 * Fz;
    fs[9] = 4.0 * u * Fx - 2.0 * v * Fy - 2.0 * w * Fz;
    fs[10] = -2.0 * u * Fx + v * Fy + w * Fz;
    fs[11] = 2.0 * v * Fy - 2.0 * w * Fz;
    fs[12] = -v * Fy + w * Fz;
    fs[13] = u * Fy + v * Fx;
    fs[14] = v * Fz + w * Fy;
    fs[15] = u * Fz + w * Fx;
    fs[16] = 0.0; fs[17] = 0.0; fs[18] = 0.0;
    for (int a = 0; a < Q; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]) + (1.0 - 0.5 * s[a]) * fs[a];   /* :767-769 */
    /* inverse transform, :771-856 */
    fp[0] = mp[0] / 19.0 - 5.0 / 399.0 * mp[1] + mp[2] / 21.0;
    fp[1] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[3] - mp[4]) * 0.1 + (mp[9] - mp[10]) / 18.0;
    fp[2] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[3] - mp[4]) * 0.1 + (mp[9] - mp[10]) / 18.0;
    fp[3] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[5] - mp[6]) * 0.1 - (mp[9] - mp[10]) / 36.0
          + (mp[11] - mp[12]) / 12.0;
    fp[4] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[5] - mp[6]) * 0.1 - (mp[9] - mp[10]) / 36.0
          + (mp[11] - mp[12]) / 12.0;
    fp[5] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[7] - mp[8]) * 0.1 - (mp[9] - mp[10]) / 36.0
          - (mp[11] - mp[12]) / 12.0;
    fp[6] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[7] - mp[8]) * 0.1 - (mp[9] - mp[10]) / 36.0
          - (mp[11] - mp[12]) / 12.0;
    fp[7] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
          + 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[5] + mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0
          + mp[13] * 0.25 + (mp[16] - mp[17]) * 0.125;
    fp[8] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
          - 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[5] - mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0
          - mp[13] * 0.25 - (mp[16] + mp[17]) * 0.125;
    fp[9] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
          + 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[5] - mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0
          - mp[13] * 0.25 + (mp[16] + mp[17]) * 0.125;
    fp[10] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           - 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[5] + mp[6])
           + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0
           + mp[13] * 0.25 - (mp[16] - mp[17]) * 0.125;
    fp[11] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           + 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[7] + mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0
           + 0.25 * mp[15] - 0.1250 * (mp[16] - mp[18]);
    fp[12] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           - 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[7] - mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0
           - 0.25 * mp[15] + 0.125 * (mp[16] + mp[18]);
    fp[13] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           + 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[7] - mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0
           - 0.25 * mp[15] - 0.125 * (mp[16] + mp[18]);
    fp[14] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           - 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[7] + mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0
           + 0.25 * mp[15] + 0.125 * (mp[16] - mp[18]);
    fp[15] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           + (4.0 * mp[5] + mp[6] + 4.0 * mp[7] + mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0
           + 0.25 * mp[14] + 0.125 * (mp[17] - mp[18]);
    fp[16] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           - (4.0 * mp[5] + mp[6] - 4.0 * mp[7] - mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0
           - 0.25 * mp[14] - 0.125 * (mp[17] + mp[18]);
    fp[17] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           + (4.0 * mp[5] + mp[6] - 4.0 * mp[7] - mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0
           - 0.25 * mp[14] + 0.125 * (mp[17] + mp[18]);
    fp[18] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0
           - (4.0 * mp[5] + mp[6] + 4.0 * mp[7] + mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0
           + 0.25 * mp[14] - 0.125 * (mp[17] - mp[18]);
}

void th_collision(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i) {
            double Fv[3];
            th_collide_cell(&F(R, 0, i, j, k), S(R, rho, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k), S(R, w, i, j, k),
                            S(R, T, i, j, k), &w->p, &FP(R, 0, i, j, k), Fv);
            S(R, Fx, i, j, k) = Fv[0]; S(R, Fy, i, j, k) = Fv[1]; S(R, Fz, i, j, k) = Fv[2];
        }
    }
}

/* ---- f_message_passing_sendrecv(), B3:1287-1415 (same messages as the lid driver) -------------------- */
static void copy_face_f(th_world *w, int dir, const int pops[5]) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *Sx = &w->r[r];
        int d = Sx->nbr_surface[dir];
        if (d < 0) continue;
        th_rank *D = &w->r[d];
        for (int q = 0; q < 5; ++q) {
            int a = pops[q];
            switch (dir) {
            case 1: for (int k = 1; k <= Sx->nz; ++k) for (int j = 1; j <= Sx->ny; ++j) FP(D, a, 0, j, k) = FP(Sx, a, Sx->nx, j, k); break;
            case 2: for (int k = 1; k <= Sx->nz; ++k) for (int j = 1; j <= Sx->ny; ++j) FP(D, a, D->nx + 1, j, k) = FP(Sx, a, 1, j, k); break;
            case 3: for (int k = 1; k <= Sx->nz; ++k) for (int i = 1; i <= Sx->nx; ++i) FP(D, a, i, 0, k) = FP(Sx, a, i, Sx->ny, k); break;
            case 4: for (int k = 1; k <= Sx->nz; ++k) for (int i = 1; i <= Sx->nx; ++i) FP(D, a, i, D->ny + 1, k) = FP(Sx, a, i, 1, k); break;
            case 5: for (int j = 1; j <= Sx->ny; ++j) for (int i = 1; i <= Sx->nx; ++i) FP(D, a, i, j, 0) = FP(Sx, a, i, j, Sx->nz); break;
            case 6: for (int j = 1; j <= Sx->ny; ++j) for (int i = 1; i <= Sx->nx; ++i) FP(D, a, i, j, D->nz + 1) = FP(Sx, a, i, j, 1); break;
            }
        }
    }
}
static void copy_edge_f(th_world *w, int a) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *Sx = &w->r[r];
        int d = Sx->nbr_line[a];
        if (d < 0) continue;
        th_rank *D = &w->r[d];
        int e[3] = {ex[a], ey[a], ez[a]};
        int sn[3] = {Sx->nx, Sx->ny, Sx->nz}, dn[3] = {D->nx, D->ny, D->nz};
        int run = (e[0] == 0) ? 0 : (e[1] == 0) ? 1 : 2;
        for (int t = 1; t <= sn[run]; ++t) {
            int si[3], di[3];
            for (int q = 0; q < 3; ++q) {
                if (q == run) { si[q] = t; di[q] = t; }
                else if (e[q] > 0) { si[q] = sn[q]; di[q] = 0; }
                else { si[q] = 1; di[q] = dn[q] + 1; }
            }
            FP(D, a, di[0], di[1], di[2]) = FP(Sx, a, si[0], si[1], si[2]);
        }
    }
}
void th_exchange_f(th_world *w) {
    static const int px[5] = {1, 7, 9, 11, 13}, mx[5] = {2, 8, 10, 12, 14};
    static const int py[5] = {3, 7, 8, 15, 17}, my[5] = {4, 9, 10, 16, 18};
    static const int pz[5] = {5, 11, 12, 15, 16}, mz[5] = {6, 13, 14, 17, 18};
    copy_face_f(w, 1, px); copy_face_f(w, 2, mx); copy_face_f(w, 3, py); copy_face_f(w, 4, my);
    copy_face_f(w, 5, pz); copy_face_f(w, 6, mz);
    static const int order[12] = {7, 10, 9, 8, 11, 14, 13, 12, 15, 18, 17, 16};
    for (int q = 0; q < 12; ++q) copy_edge_f(w, order[q]);
}

/* ---- g_message_passing_sendrecv(), B3:1421-1468: one population per face, no edges ------------------- */
void th_exchange_g(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *Sx = &w->r[r];
        for (int dir = 1; dir <= 6; ++dir) {
            int d = Sx->nbr_surface[dir];
            if (d < 0) continue;
            th_rank *D = &w->r[d];
            const int a = dir;      /* +x:1 -x:2 +y:3 -y:4 +z:5 -z:6 */
            switch (dir) {
            case 1: for (int k = 1; k <= Sx->nz; ++k) for (int j = 1; j <= Sx->ny; ++j) GP(D, a, 0, j, k) = GP(Sx, a, Sx->nx, j, k); break;
            case 2: for (int k = 1; k <= Sx->nz; ++k) for (int j = 1; j <= Sx->ny; ++j) GP(D, a, D->nx + 1, j, k) = GP(Sx, a, 1, j, k); break;
            case 3: for (int k = 1; k <= Sx->nz; ++k) for (int i = 1; i <= Sx->nx; ++i) GP(D, a, i, 0, k) = GP(Sx, a, i, Sx->ny, k); break;
            case 4: for (int k = 1; k <= Sx->nz; ++k) for (int i = 1; i <= Sx->nx; ++i) GP(D, a, i, D->ny + 1, k) = GP(Sx, a, i, 1, k); break;
            case 5: for (int j = 1; j <= Sx->ny; ++j) for (int i = 1; i <= Sx->nx; ++i) GP(D, a, i, j, 0) = GP(Sx, a, i, j, Sx->nz); break;
            case 6: for (int j = 1; j <= Sx->ny; ++j) for (int i = 1; i <= Sx->nx; ++i) GP(D, a, i, j, D->nz + 1) = GP(Sx, a, i, j, 1); break;
            }
        }
    }
}

/* ---- streaming(), B3:868-892 / streamingT(), B3:1075-1098 ----------------------------------------------- */
void th_streaming(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i)
            for (int a = 0; a < Q; ++a) F(R, a, i, j, k) = FP(R, a, i - ex[a], j - ey[a], k - ez[a]);
    }
}
void th_streamingT(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i)
            for (int a = 0; a < QT; ++a) G(R, a, i, j, k) = GP(R, a, i - ex[a], j - ey[a], k - ez[a]);
    }
}

/* ---- bounceback(), B3:895-983: no-slip half-way bounce-back on all six walls (noslipWalls) ------------ */
void th_bounceback(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny, nz = R->nz;
        if (R->coords[0] == 0)
            for (int k = 1; k <= nz; ++k) for (int j = 1; j <= ny; ++j) {
                F(R, 1, 1, j, k) = FP(R, 2, 1, j, k);   F(R, 7, 1, j, k) = FP(R, 10, 1, j, k);
                F(R, 9, 1, j, k) = FP(R, 8, 1, j, k);   F(R, 11, 1, j, k) = FP(R, 14, 1, j, k);
                F(R, 13, 1, j, k) = FP(R, 12, 1, j, k);
            }
        if (R->coords[0] == w->dims[0] - 1)
            for (int k = 1; k <= nz; ++k) for (int j = 1; j <= ny; ++j) {
                F(R, 2, nx, j, k) = FP(R, 1, nx, j, k);  F(R, 8, nx, j, k) = FP(R, 9, nx, j, k);
                F(R, 10, nx, j, k) = FP(R, 7, nx, j, k); F(R, 12, nx, j, k) = FP(R, 13, nx, j, k);
                F(R, 14, nx, j, k) = FP(R, 11, nx, j, k);
            }
        if (R->coords[1] == 0)
            for (int k = 1; k <= nz; ++k) for (int i = 1; i <= nx; ++i) {
                F(R, 3, i, 1, k) = FP(R, 4, i, 1, k);   F(R, 7, i, 1, k) = FP(R, 10, i, 1, k);
                F(R, 8, i, 1, k) = FP(R, 9, i, 1, k);   F(R, 15, i, 1, k) = FP(R, 18, i, 1, k);
                F(R, 17, i, 1, k) = FP(R, 16, i, 1, k);
            }
        if (R->coords[1] == w->dims[1] - 1)
            for (int k = 1; k <= nz; ++k) for (int i = 1; i <= nx; ++i) {
                F(R, 4, i, ny, k) = FP(R, 3, i, ny, k);  F(R, 9, i, ny, k) = FP(R, 8, i, ny, k);
                F(R, 10, i, ny, k) = FP(R, 7, i, ny, k); F(R, 16, i, ny, k) = FP(R, 17, i, ny, k);
                F(R, 18, i, ny, k) = FP(R, 15, i, ny, k);
            }
        if (R->coords[2] == 0)
            for (int j = 1; j <= ny; ++j) for (int i = 1; i <= nx; ++i) {
                F(R, 5, i, j, 1) = FP(R, 6, i, j, 1);   F(R, 11, i, j, 1) = FP(R, 14, i, j, 1);
                F(R, 12, i, j, 1) = FP(R, 13, i, j, 1); F(R, 15, i, j, 1) = FP(R, 18, i, j, 1);
                F(R, 16, i, j, 1) = FP(R, 17, i, j, 1);
            }
        if (R->coords[2] == w->dims[2] - 1)
            for (int j = 1; j <= ny; ++j) for (int i = 1; i <= nx; ++i) {
                F(R, 6, i, j, nz) = FP(R, 5, i, j, nz);   F(R, 13, i, j, nz) = FP(R, 12, i, j, nz);
                F(R, 14, i, j, nz) = FP(R, 11, i, j, nz); F(R, 17, i, j, nz) = FP(R, 16, i, j, nz);
                F(R, 18, i, j, nz) = FP(R, 15, i, j, nz);
            }
    }
}

/* ---- collisionT(), B3:1014-1070: one cell --------------------------------------------------------------- */
void th_collideT_cell(const double *g, double u, double v, double w, double T, const th_params *p, double *gp) {
    double n[QT], neq[QT], q[QT], np_[QT];
    n[0] = g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];
    n[1] = g[1] - g[2];
    n[2] = g[3] - g[4];
    n[3] = g[5] - g[6];
    n[4] = -6.0 * g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];
    n[5] = 2.0 * g[1] + 2.0 * g[2] - g[3] - g[4] - g[5] - g[6];
    n[6] = g[3] + g[4] - g[5] - g[6];
    neq[0] = T; neq[1] = T * u; neq[2] = T * v; neq[3] = T * w; neq[4] = T * p->paraA; neq[5] = 0.0; neq[6] = 0.0;
    q[0] = 0.0; q[1] = p->Qd; q[2] = p->Qd; q[3] = p->Qd; q[4] = p->Qnu; q[5] = p->Qnu; q[6] = p->Qnu;
    for (int a = 0; a < QT; ++a) np_[a] = n[a] - q[a] * (n[a] - neq[a]);
    gp[0] = np_[0] / 7.0 - np_[4] / 7.0;
    gp[1] = np_[0] / 7.0 + 0.5 * np_[1] + np_[4] / 42.0 + np_[5] / 6.0;
    gp[2] = np_[0] / 7.0 - 0.5 * np_[1] + np_[4] / 42.0 + np_[5] / 6.0;
    gp[3] = np_[0] / 7.0 + 0.5 * np_[2] + np_[4] / 42.0 - np_[5] / 12.0 + 0.25 * np_[6];
    gp[4] = np_[0] / 7.0 - 0.5 * np_[2] + np_[4] / 42.0 - np_[5] / 12.0 + 0.25 * np_[6];
    gp[5] = np_[0] / 7.0 + 0.5 * np_[3] + np_[4] / 42.0 - np_[5] / 12.0 - 0.25 * np_[6];
    gp[6] = np_[0] / 7.0 - 0.5 * np_[3] + np_[4] / 42.0 - np_[5] / 12.0 - 0.25 * np_[6];
}

void th_collisionT(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i)
            th_collideT_cell(&G(R, 0, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k), S(R, w, i, j, k), S(R, T, i, j, k),
                             &w->p, &GP(R, 0, i, j, k));
    }
}

/* ---- bouncebackT(), B3:1100-1210: adiabatic g_a = g_post_opp, constant T g_a = -g_post_opp + (6+paraA)/21*Tw */
void th_bouncebackT(th_world *w) {
    const double paraA = w->p.paraA;
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        const int n[3] = {R->nx, R->ny, R->nz};
        for (int face = 0; face < 6; ++face) {
            const int axis = face >> 1, minus = face & 1;
            if (minus ? (R->coords[axis] != 0) : (R->coords[axis] != w->dims[axis] - 1)) continue;
            const int fix = minus ? 1 : n[axis];
            const int a = minus ? 2 * axis + 1 : 2 * axis + 2;      /* population leaving the wall */
            const int o = minus ? 2 * axis + 2 : 2 * axis + 1;      /* its opposite */
            const int kind = w->bcT[face];
            const double Tw = kind == TH_CONST_HOT ? w->p.Thot : w->p.Tcold;
            const int n1 = axis == 0 ? n[1] : n[0], n2 = axis == 2 ? n[1] : n[2];
            for (int t2 = 1; t2 <= n2; ++t2)
            for (int t1 = 1; t1 <= n1; ++t1) {
                const int i = axis == 0 ? fix : t1, j = axis == 1 ? fix : (axis == 0 ? t1 : t2), k = axis == 2 ? fix : t2;
                if (kind == TH_ADIABATIC) G(R, a, i, j, k) = GP(R, o, i, j, k);
                else G(R, a, i, j, k) = -GP(R, o, i, j, k) + (6.0 + paraA) / 21.0 * Tw;
            }
        }
    }
}

/* ---- macro(), B3:986-1010 / macroT(), B3:1215-1232 ------------------------------------------------------- */
void th_macro_cell(const double *f, double Fx, double Fy, double Fz, double *out /* rho,u,v,w */) {
    double rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6]
               + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    out[0] = rho;
    out[1] = (f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14] + 0.5 * Fx) / rho;
    out[2] = (f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18] + 0.5 * Fy) / rho;
    out[3] = (f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18] + 0.5 * Fz) / rho;
}
void th_macro(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i) {
            double o[4];
            th_macro_cell(&F(R, 0, i, j, k), S(R, Fx, i, j, k), S(R, Fy, i, j, k), S(R, Fz, i, j, k), o);
            S(R, rho, i, j, k) = o[0]; S(R, u, i, j, k) = o[1]; S(R, v, i, j, k) = o[2]; S(R, w, i, j, k) = o[3];
        }
    }
}
void th_macroT(th_world *w) {
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i) {
            const double *g = &G(R, 0, i, j, k);
            S(R, T, i, j, k) = g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];
        }
    }
}

/* ---- check(), B3:1236-1283: errorU WITH the w term (unlike the lid driver), errorT = sum|dT| / sum|T| ---- */
void th_check(th_world *w, double *errorU, double *errorT) {
    double t1 = 0.0, t2 = 0.0, t5 = 0.0, t6 = 0.0;
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        double e1 = 0.0, e2 = 0.0, e5 = 0.0, e6 = 0.0;
        size_t n = (size_t)R->nx * R->ny * R->nz;
        for (size_t q = 0; q < n; ++q) {
            double u = R->u[q], v = R->v[q], ww = R->w[q], T = R->T[q];
            e1 = e1 + (u - R->up[q]) * (u - R->up[q]) + (v - R->vp[q]) * (v - R->vp[q]) + (ww - R->wp[q]) * (ww - R->wp[q]);
            e2 = e2 + u * u + v * v + ww * ww;
            e5 = e5 + fabs(T - R->Tp[q]);
            e6 = e6 + fabs(T);
            R->up[q] = u; R->vp[q] = v; R->wp[q] = ww; R->Tp[q] = T;
        }
        t1 += e1; t2 += e2; t5 += e5; t6 += e6;
    }
    w->errorU = sqrt(t1) / sqrt(t2);
    w->errorT = t5 / t6;
    *errorU = w->errorU; *errorT = w->errorT;
}

/* ---- calNuRe(), B3/mpi_blocked/RaNu.F90:13-47: loop-order sums (k, j, i) per rank, rank-ordered sum over the ranks (the
 * Allreduce the MPI driver would need), averages over the global box; viscosity, diffusivity from module.F90:38-39 ---- */
void th_calNuRe(th_world *w, double prandtl, double *NuVolAvg, double *ReVolAvg) {
    double t1 = 0.0, t2 = 0.0;
    for (int r = 0; r < w->np; ++r) {
        th_rank *R = &w->r[r];
        double NuVolAvg_temp = 0.0, ReVolAvg_temp = 0.0;
        size_t n = (size_t)R->nx * R->ny * R->nz;
        for (size_t q = 0; q < n; ++q) NuVolAvg_temp = NuVolAvg_temp + R->w[q] * R->T[q];
        for (size_t q = 0; q < n; ++q)
            ReVolAvg_temp = ReVolAvg_temp + (R->u[q] * R->u[q] + R->v[q] * R->v[q] + R->w[q] * R->w[q]);
        t1 += NuVolAvg_temp; t2 += ReVolAvg_temp;
    }
    const double viscosity = (w->p.tauf - 0.5) / 3.0, diffusivity = viscosity / prandtl;
    const double ncell = (double)((long long)w->total[0] * w->total[1] * w->total[2]), nz = (double)w->total[2];
    *NuVolAvg = t1 / ncell * nz / diffusivity + 1.0;
    *ReVolAvg = sqrt(t2 / ncell) * nz / viscosity;
}

/* n iterations of the driver loop body, B3:222-248 */
void th_step(th_world *w, int n) {
    for (int s = 0; s < n; ++s) {
        w->itc += 1;
        th_collision(w);
        th_exchange_f(w);
        th_streaming(w);
        th_bounceback(w);
        th_collisionT(w);
        th_exchange_g(w);
        th_streamingT(w);
        th_bouncebackT(w);
        th_macro(w);
        th_macroT(w);
    }
}
