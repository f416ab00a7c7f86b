/*
 * oracle/thermal2d.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's 2-D double-distribution thermal
 * lattice Boltzmann driver (D2Q9 MRT flow with Boussinesq forcing + D2Q5 MRT temperature):
 *   B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/  (Fortran + MPI, 2-D Cartesian blocks, 201 x 201, Ra = 1e7)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product never does.
 *
 * PARITY PIN: B2 is Fortran + MPI, which this image cannot build.  The restatement is pinned to the reference's own
 * source TEXT: tests/golden/make_golden_thermal2d.py machine-evaluates collision(), collisionT(), macro(), macroT(),
 * bounceback(), bouncebackT(), streaming(), streamingT(), the initial() loops and check()'s sums from B2's files
 * (fortran_eval.py) on seeded inputs, and tests/test_oracle_thermal2d.py requires this file to reproduce those numbers
 * bit for bit; on top: the reference's seq == MPI contract (P emulated ranks == 1 rank, bit for bit) and analytic pins.
 * Whole run: the sequential side-heated program seq/steady.F90 is evaluated from its text on 9 x 7 -- parameters, initial() and
 * its loop for 1, 2, 20, 25 iterations with check() (make_golden_thermal2d_seq_run.py) -- and this file (variant T2_MPI,
 * side-heated walls) reproduces its f, g, rho, u, v, T, Fx, Fy bit for bit on 1..6 emulated ranks.  The OpenACC program
 * seq/bouyancy2d_acc.F90 is evaluated the same way (periodic vertical walls, Rayleigh-Benard plates, lengthUnit = nx) and
 * variant T2_ACC reproduces that run bit for bit on 1..3 ranks stacked along y (ref_fortran_thermal2d_acc_run.npz).
 *
 * Layout is B2's: column-major, population index fastest: f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), g(0:4,nx,ny),
 * g_post(0:4,0:nx+1,0:ny+1), rho,u,v,T,up,vp,Tp,Fx,Fy(nx,ny)  (initial.F90:177-197).
 * Left-to-right evaluation, true divisions, -ffp-contract=off: every operation is one IEEE fp64 rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q9 9
#define Q5 5
enum { T2_ADIABATIC = 0, T2_CONST_HOT = 1, T2_CONST_COLD = 2, T2_PERIODIC = 3 };
enum { T2_MPI = 0, T2_ACC = 1 };

/* module.F90:106-109 */
static const int ex[Q9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int ey[Q9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

typedef struct t2_params {
    double Rayleigh, Prandtl, Mach, Thot, Tcold, Tref, rho0;                         /* inputs  module.F90:31-33,67-68 */
    double lengthUnit, tauf, viscosity, diffusivity, paraA, gBeta, Snu, Sq, Qd, Qnu; /* derived module.F90:29,69-81    */
} t2_params;

typedef struct t2_rank {
    int nx, ny, coords[2], start[2];
    int nbr[4];          /* right(+x), left(-x), top(+y), bottom(-y); -1 = MPI_PROC_NULL     main.F90:41-42 */
    int cnr[4];          /* top_right(5), top_left(6), bottom_left(7), bottom_right(8)       MPI_Cart_find_corners */
    double *f, *f_post, *g, *g_post, *rho, *u, *v, *T, *up, *vp, *Tp, *Fx, *Fy;
} t2_rank;

typedef struct t2_world {
    int total[2], dims[2], np, itc;
    int bcT[4];          /* +x (right), -x (left), +y (top), -y (bottom) : T2_*   macros.F90:16-27 */
    int variant;         /* T2_MPI: mpi_blocked/evolution_f.F90 ; T2_ACC: seq/bouyancy2d_acc.F90 (f_post(0) rounded term by term) */
    int periodic_x;      /* VerticalWallsPeriodicalU + VerticalWallsPeriodicalT (acc:15,22): needs dims[0] == 1 */
    /* the sheared Rayleigh-Benard programs (RB2 = seq/R_B_2d.F90; seq/bouyancy2d_omp.F90): walls that move along themselves.
     * Uwall = TopLeft, TopRight, BottomLeft, BottomRight (u of the horizontal walls, left / right half: i <= nxHalf or not),
     * LeftTop, LeftBottom, RightTop, RightBottom (v of the vertical walls, j <= nyHalf = Bottom)            RB2:87,118-120 */
    double Uwall[8];
    int moving;          /* any Uwall != 0 */
    int cornersT;        /* 1 = RB2:1086-1106: both populations of a corner cell take the plate's constant-temperature rule */
    t2_params p;
    double errorU, errorT;
    t2_rank *r;
} t2_world;

#define F(R, a, i, j) ((R)->f[(a) + Q9 * ((size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1))])
#define FP(R, a, i, j) ((R)->f_post[(a) + Q9 * ((size_t)(i) + (size_t)((R)->nx + 2) * (size_t)(j))])
#define G(R, a, i, j) ((R)->g[(a) + Q5 * ((size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1))])
#define GP(R, a, i, j) ((R)->g_post[(a) + Q5 * ((size_t)(i) + (size_t)((R)->nx + 2) * (size_t)(j))])
#define S(R, A, i, j) ((R)->A[(size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1)])

/* module.F90:29,69-81.  lengthUnit = dble(total_ny); Rayleigh is the single-precision literal 1e7 (exact). */
void t2_derive_params(t2_params *p, int total_ny) {
    if (!(p->lengthUnit > 0.0)) p->lengthUnit = (double)total_ny;        /* acc:57 uses dble(nx) instead: the caller sets it */
    p->tauf = 0.5 + p->Mach * p->lengthUnit * sqrt(3.0 * p->Prandtl / p->Rayleigh);
    p->viscosity = (p->tauf - 0.5) / 3.0;
    p->diffusivity = p->viscosity / p->Prandtl;
    p->paraA = 20.0 * sqrt(3.0) * p->diffusivity - 4.0;
    const double gBeta1 = p->Rayleigh * p->viscosity * p->diffusivity / p->lengthUnit;
    p->gBeta = gBeta1 / p->lengthUnit / p->lengthUnit;
    p->Snu = 1.0 / p->tauf;
    p->Sq = 8.0 * (2.0 * p->tauf - 1.0) / (8.0 * p->tauf - 1.0);
    p->Qd = 3.0 - sqrt(3.0);
    p->Qnu = 4.0 * sqrt(3.0) - 6.0;
}

/* MPI_Dims_create(np, 2, dims) with dims = 0 (main.F90:22): balanced, non-increasing */
static void dims_create2(int np, int dims[2]) {
    int best = np;
    for (int a = 1; a <= np; ++a)
        if (np % a == 0 && a >= np / a && a < best) best = a;
    dims[0] = best; dims[1] = np / best;
}
static void decompose_1d(int total_n, int rank, int np, int *local_n, int *start) {   /* main.F90:208-225 */
    int n = total_n / np, m = total_n % np;
    *local_n = n + (rank < m ? 1 : 0);
    *start = rank * n + (rank < m ? rank : m);
}
static int cart_rank(const int dims[2], int c0, int c1) {
    if (c0 < 0 || c0 >= dims[0] || c1 < 0 || c1 >= dims[1]) return -1;
    return c0 * dims[1] + c1;
}

/* par7 = Rayleigh, Prandtl, Mach, Thot, Tcold, Tref, rho0;  bcT_or_null = NULL: the shipped side-heated cell
 * (macros.F90:24-27: vertical walls constant T, hot on the left; horizontal walls adiabatic) */
t2_world *t2_world_create(int tnx, int tny, int np, const int *dims_or_null, const double *par7, const int *bcT_or_null) {
    static const int shipped[4] = {T2_CONST_COLD, T2_CONST_HOT, T2_ADIABATIC, T2_ADIABATIC};
    t2_world *w = (t2_world *)calloc(1, sizeof(t2_world));
    w->total[0] = tnx; w->total[1] = tny; w->np = np;
    if (dims_or_null && dims_or_null[0] > 0) memcpy(w->dims, dims_or_null, 2 * sizeof(int));
    else dims_create2(np, w->dims);
    memcpy(w->bcT, bcT_or_null ? bcT_or_null : shipped, sizeof w->bcT);
    w->p.Rayleigh = par7[0]; w->p.Prandtl = par7[1]; w->p.Mach = par7[2]; w->p.Thot = par7[3]; w->p.Tcold = par7[4];
    w->p.Tref = par7[5]; w->p.rho0 = par7[6];
    w->p.lengthUnit = 0.0;
    w->periodic_x = w->bcT[0] == T2_PERIODIC || w->bcT[1] == T2_PERIODIC;
    t2_derive_params(&w->p, tny);
    w->r = (t2_rank *)calloc((size_t)np, sizeof(t2_rank));
    for (int c0 = 0; c0 < w->dims[0]; ++c0)
        for (int c1 = 0; c1 < w->dims[1]; ++c1) {
            t2_rank *R = &w->r[cart_rank(w->dims, c0, c1)];
            R->coords[0] = c0; R->coords[1] = c1;
            decompose_1d(tnx, c0, w->dims[0], &R->nx, &R->start[0]);
            decompose_1d(tny, c1, w->dims[1], &R->ny, &R->start[1]);
            R->nbr[0] = cart_rank(w->dims, c0 + 1, c1); R->nbr[1] = cart_rank(w->dims, c0 - 1, c1);
            R->nbr[2] = cart_rank(w->dims, c0, c1 + 1); R->nbr[3] = cart_rank(w->dims, c0, c1 - 1);
            for (int a = 5; a < Q9; ++a) R->cnr[a - 5] = cart_rank(w->dims, c0 + ex[a], c1 + ey[a]);
            size_t n = (size_t)R->nx * R->ny, nh = (size_t)(R->nx + 2) * (R->ny + 2);
            R->f = (double *)calloc(Q9 * n, sizeof(double)); R->f_post = (double *)calloc(Q9 * nh, sizeof(double));
            R->g = (double *)calloc(Q5 * n, sizeof(double)); R->g_post = (double *)calloc(Q5 * nh, sizeof(double));
            double **fld[] = {&R->rho, &R->u, &R->v, &R->T, &R->up, &R->vp, &R->Tp, &R->Fx, &R->Fy};
            for (size_t q = 0; q < sizeof fld / sizeof fld[0]; ++q) *fld[q] = (double *)calloc(n, sizeof(double));
        }
    return w;
}
void t2_world_destroy(t2_world *w) {
    if (!w) return;
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        double *all[] = {R->f, R->f_post, R->g, R->g_post, R->rho, R->u, R->v, R->T, R->up, R->vp, R->Tp, R->Fx, R->Fy};
        for (size_t q = 0; q < sizeof all / sizeof all[0]; ++q) free(all[q]);
    }
    free(w->r); free(w);
}
/* the OpenACC program's flavour: variant = T2_ACC, lengthUnit = dble(nx) (acc:57); re-derives the parameters */
void t2_world_set_variant(t2_world *w, int variant, double lengthUnit_or_0) {
    w->variant = variant;
    w->p.lengthUnit = lengthUnit_or_0;
    t2_derive_params(&w->p, w->total[1]);
}
/* moving walls and the corner rule of the sheared Rayleigh-Benard programs (RB2:118-120, :1086-1106) */
void t2_world_set_walls(t2_world *w, const double *Uwall8, int cornersT) {
    w->moving = 0;
    for (int q = 0; q < 8; ++q) { w->Uwall[q] = Uwall8 ? Uwall8[q] : 0.0; w->moving |= w->Uwall[q] != 0.0; }
    w->cornersT = cornersT;
}
/* RB2:87; velocity of the horizontal wall (top = 1 / bottom = 0) above or below global column gi, of the vertical wall
 * (right = 1 / left = 0) beside global row gj */
static double wall_u(const t2_world *w, int top, int gi) {
    const int left = gi <= (w->total[0] - 1) / 2 + 1;
    return w->Uwall[top ? (left ? 0 : 1) : (left ? 2 : 3)];
}
static double wall_v(const t2_world *w, int right, int gj) {
    const int bottom = gj <= (w->total[1] - 1) / 2 + 1;
    return w->Uwall[right ? (bottom ? 7 : 6) : (bottom ? 5 : 4)];
}
void t2_world_info(t2_world *w, int dims[2], t2_params *p, int bcT[4]) {
    dims[0] = w->dims[0]; dims[1] = w->dims[1];
    *p = w->p;
    memcpy(bcT, w->bcT, sizeof w->bcT);
}
void t2_rank_info(t2_world *w, int r, int info[14]) {
    t2_rank *R = &w->r[r];
    info[0] = R->nx; info[1] = R->ny; info[2] = R->coords[0]; info[3] = R->coords[1]; info[4] = R->start[0]; info[5] = R->start[1];
    memcpy(info + 6, R->nbr, sizeof R->nbr); memcpy(info + 10, R->cnr, sizeof R->cnr);
}
double *t2_rank_ptr(t2_world *w, int r, int which) {
    t2_rank *R = &w->r[r];
    double *p[] = {R->f, R->f_post, R->g, R->g_post, R->rho, R->u, R->v, R->T, R->up, R->vp, R->Tp, R->Fx, R->Fy};
    return p[which];
}

/* initial(): initial.F90:199-212 (weights), :245-272 (fields), :276-288 (populations), :326-335 */
void t2_initial(t2_world *w) {
    const t2_params *p = &w->p;
    double omega[Q9], omegaT[Q5];
    omega[0] = 4.0 / 9.0;
    for (int a = 1; a <= 4; ++a) omega[a] = 1.0 / 9.0;
    for (int a = 5; a <= 8; ++a) omega[a] = 1.0 / 36.0;
    omegaT[0] = (1.0 - p->paraA) / 5.0;
    for (int a = 1; a <= 4; ++a) omegaT[a] = (p->paraA + 4.0) / 20.0;
    w->itc = 0; w->errorU = 100.0; w->errorT = 100.0;
    const int isT[4] = {w->bcT[0] == T2_CONST_HOT || w->bcT[0] == T2_CONST_COLD, w->bcT[1] == T2_CONST_HOT || w->bcT[1] == T2_CONST_COLD,
                        w->bcT[2] == T2_CONST_HOT || w->bcT[2] == T2_CONST_COLD, w->bcT[3] == T2_CONST_HOT || w->bcT[3] == T2_CONST_COLD};
    const int vertT = isT[0] || isT[1];                                            /* #ifdef VerticalWallsConstT   */
    const int horT = isT[2] || isT[3];                                             /* #ifdef HorizontalWallsConstT */
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                S(R, rho, i, j) = p->rho0; S(R, u, i, j) = 0.0; S(R, v, i, j) = 0.0; S(R, T, i, j) = 0.0;
                S(R, up, i, j) = 0.0; S(R, vp, i, j) = 0.0; S(R, Tp, i, j) = 0.0;
                if (vertT) S(R, T, i, j) = (double)(R->start[0] + i - 1) / (double)(w->total[0] - 1) * (p->Tcold - p->Thot) + p->Thot;
                if (horT) S(R, T, i, j) = (double)(R->start[1] + j - 1) / (double)(w->total[1] - 1) * (p->Tcold - p->Thot) + p->Thot;
                if (w->moving) {                                                   /* RB2:466-481 */
                    const int gi = R->start[0] + i, gj = R->start[1] + j;
                    if (gj == w->total[1]) S(R, u, i, j) = wall_u(w, 1, gi);
                    if (gj == 1) S(R, u, i, j) = wall_u(w, 0, gi);
                    if (gj >= 2 && gj <= w->total[1] - 1) {
                        if (gi == 1) S(R, v, i, j) = wall_v(w, 0, gj);
                        if (gi == w->total[0]) S(R, v, i, j) = wall_v(w, 1, gj);
                    }
                }
            }
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                double un[Q9];
                double us2 = S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j);
                for (int a = 0; a < Q9; ++a) {
                    un[a] = S(R, u, i, j) * (double)ex[a] + S(R, v, i, j) * (double)ey[a];
                    F(R, a, i, j) = S(R, rho, i, j) * omega[a] * (1.0 + 3.0 * un[a] + 4.5 * un[a] * un[a] - 1.5 * us2);
                }
                for (int a = 0; a < Q5; ++a) {
                    un[a] = S(R, u, i, j) * (double)ex[a] + S(R, v, i, j) * (double)ey[a];
                    G(R, a, i, j) = S(R, T, i, j) * omegaT[a] * (1.0 + 10.0 / (4.0 + p->paraA) * un[a]);
                }
            }
        memset(R->f_post, 0, sizeof(double) * Q9 * (size_t)(R->nx + 2) * (R->ny + 2));
        memset(R->g_post, 0, sizeof(double) * Q5 * (size_t)(R->nx + 2) * (R->ny + 2));
    }
}

/* collision() of one cell: evolution_f.F90:15-78.  out: fp[9], FxFy[2] */
void t2_collide_cell_v(int variant, const t2_params *p, const double *f, double rho, double u, double v, double T, double *fp, double *FxFy);
void t2_collide_cell(const t2_params *p, const double *f, double rho, double u, double v, double T, double *fp, double *FxFy) {
    t2_collide_cell_v(T2_MPI, p, f, rho, u, v, T, fp, FxFy);
}
/* variant T2_ACC: seq/bouyancy2d_acc.F90:628-694 -- identical except f_post(0) = m0/9 - m1/9 + m2/9 (acc:679) */
void t2_collide_cell_v(int variant, const t2_params *p, const double *f, double rho, double u, double v, double T, double *fp, double *FxFy) {
    double m[Q9], meq[Q9], mp[Q9], s[Q9], fs[Q9];
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * (f[5] + f[6] + f[7] + f[8]);
    m[2] = 4.0 * f[0] - 2.0 * (f[1] + f[2] + f[3] + f[4]) + f[5] + f[6] + f[7] + f[8];
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
    meq[4] = -rho * u;
    meq[5] = rho * v;
    meq[6] = -rho * v;
    meq[7] = rho * (u * u - v * v);
    meq[8] = rho * (u * v);
    s[0] = 0.0; s[1] = p->Snu; s[2] = p->Snu; s[3] = 0.0; s[4] = p->Sq; s[5] = 0.0; s[6] = p->Sq; s[7] = p->Snu; s[8] = p->Snu;
    const double Fx = 0.0;                                     /* :45 */
    const double Fy = rho * p->gBeta * (T - p->Tref);          /* :46 */
    fs[0] = 0.0;
    fs[1] = (6.0 - 3.0 * s[1]) * (u * Fx + v * Fy);
    fs[2] = -(6.0 - 3.0 * s[2]) * (u * Fx + v * Fy);
    fs[3] = (1.0 - 0.5 * s[3]) * Fx;
    fs[4] = -(1.0 - 0.5 * s[4]) * Fx;
    fs[5] = (1.0 - 0.5 * s[5]) * Fy;
    fs[6] = -(1.0 - 0.5 * s[6]) * Fy;
    fs[7] = (2.0 - s[7]) * (u * Fx - v * Fy);
    fs[8] = (1.0 - 0.5 * s[8]) * (u * Fy + v * Fx);
    for (int a = 0; a < Q9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]) + fs[a];
    fp[0] = variant == T2_ACC ? mp[0] / 9.0 - mp[1] / 9.0 + mp[2] / 9.0 : (mp[0] - mp[1] + mp[2]) / 9.0;
    fp[1] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[3] / 6.0 - mp[4] / 6.0 + mp[7] / 4.0;
    fp[2] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[5] / 6.0 - mp[6] / 6.0 - mp[7] / 4.0;
    fp[3] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[3] / 6.0 + mp[4] / 6.0 + mp[7] / 4.0;
    fp[4] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[5] / 6.0 + mp[6] / 6.0 - mp[7] / 4.0;
    fp[5] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 + mp[8] / 4.0;
    fp[6] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 - mp[8] / 4.0;
    fp[7] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 + mp[8] / 4.0;
    fp[8] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 - mp[8] / 4.0;
    FxFy[0] = Fx; FxFy[1] = Fy;
}
void t2_collision(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                double F2[2];
                t2_collide_cell_v(w->variant, &w->p, &F(R, 0, i, j), S(R, rho, i, j), S(R, u, i, j), S(R, v, i, j), S(R, T, i, j), &FP(R, 0, i, j), F2);
                S(R, Fx, i, j) = F2[0]; S(R, Fy, i, j) = F2[1];
            }
    }
}

/* collisionT() of one cell: evolution_g.F90:13-39 */
void t2_collideT_cell(const t2_params *p, const double *g, double u, double v, double T, double *gp) {
    double n[Q5], neq[Q5], np_[Q5], q[Q5];
    n[0] = g[0] + g[1] + g[2] + g[3] + g[4];
    n[1] = g[1] - g[3];
    n[2] = g[2] - g[4];
    n[3] = -4.0 * g[0] + g[1] + g[2] + g[3] + g[4];
    n[4] = g[1] - g[2] + g[3] - g[4];
    neq[0] = T;
    neq[1] = T * u;
    neq[2] = T * v;
    neq[3] = T * p->paraA;
    neq[4] = 0.0;
    q[0] = 0.0; q[1] = p->Qd; q[2] = p->Qd; q[3] = p->Qnu; q[4] = p->Qnu;
    for (int a = 0; a < Q5; ++a) np_[a] = n[a] - q[a] * (n[a] - neq[a]);
    gp[0] = 0.2 * np_[0] - 0.2 * np_[3];
    gp[1] = 0.2 * np_[0] + 0.5 * np_[1] + 0.05 * np_[3] + 0.25 * np_[4];
    gp[2] = 0.2 * np_[0] + 0.5 * np_[2] + 0.05 * np_[3] - 0.25 * np_[4];
    gp[3] = 0.2 * np_[0] - 0.5 * np_[1] + 0.05 * np_[3] + 0.25 * np_[4];
    gp[4] = 0.2 * np_[0] - 0.5 * np_[2] + 0.05 * np_[3] - 0.25 * np_[4];
}
void t2_collisionT(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                t2_collideT_cell(&w->p, &G(R, 0, i, j), S(R, u, i, j), S(R, v, i, j), S(R, T, i, j), &GP(R, 0, i, j));
    }
}

/* message_passing_f(): message_exchange.F90:1-79 -- 3 populations per face over the interior range, 1 per corner */
void t2_exchange_f(t2_world *w) {
    static const int face_pops[4][3] = {{1, 5, 8}, {3, 6, 7}, {2, 5, 6}, {4, 7, 8}};   /* to right, left, top, bottom */
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        for (int face = 0; face < 4; ++face) {
            if (R->nbr[face] < 0) continue;
            t2_rank *D = &w->r[R->nbr[face]];
            for (int s = 0; s < 3; ++s) {
                int a = face_pops[face][s];
                if (face < 2) for (int j = 1; j <= R->ny; ++j) FP(D, a, face == 0 ? 0 : D->nx + 1, j) = FP(R, a, face == 0 ? R->nx : 1, j);
                else for (int i = 1; i <= R->nx; ++i) FP(D, a, i, face == 2 ? 0 : D->ny + 1) = FP(R, a, i, face == 2 ? R->ny : 1);
            }
        }
        for (int a = 5; a < Q9; ++a) {
            if (R->cnr[a - 5] < 0) continue;
            t2_rank *D = &w->r[R->cnr[a - 5]];
            FP(D, a, ex[a] > 0 ? 0 : D->nx + 1, ey[a] > 0 ? 0 : D->ny + 1) = FP(R, a, ex[a] > 0 ? R->nx : 1, ey[a] > 0 ? R->ny : 1);
        }
    }
}
/* message_passing_g(): message_exchange.F90:85-118 -- the one population that crosses each face, no corners */
void t2_exchange_g(t2_world *w) {
    static const int face_pop[4] = {1, 3, 2, 4};
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        for (int face = 0; face < 4; ++face) {
            if (R->nbr[face] < 0) continue;
            t2_rank *D = &w->r[R->nbr[face]];
            int a = face_pop[face];
            if (face < 2) for (int j = 1; j <= R->ny; ++j) GP(D, a, face == 0 ? 0 : D->nx + 1, j) = GP(R, a, face == 0 ? R->nx : 1, j);
            else for (int i = 1; i <= R->nx; ++i) GP(D, a, i, face == 2 ? 0 : D->ny + 1) = GP(R, a, i, face == 2 ? R->ny : 1);
        }
    }
}

/* streaming(): evolution_f.F90:89-108 ; streamingT(): evolution_g.F90:49-68 (pull; wall halos are read as they are --
 * zero after initial() -- and overwritten by bounceback()/bouncebackT()) */
void t2_streaming(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                for (int a = 0; a < Q9; ++a) F(R, a, i, j) = FP(R, a, i - ex[a], j - ey[a]);
    }
}
void t2_streamingT(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                for (int a = 0; a < Q5; ++a) G(R, a, i, j) = GP(R, a, i - ex[a], j - ey[a]);
    }
}

/* bounceback(): evolution_f.F90:283-321 (VerticalWallsNoslip, HorizontalWallsNoslip: left, right, bottom, top) */
void t2_bounceback(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny;
        if (w->periodic_x) {   /* VerticalWallsPeriodicalU, acc:777-791: the SAME row j of the opposite side, also for the diagonals */
            for (int j = 1; j <= ny; ++j) {
                F(R, 1, 1, j) = FP(R, 1, nx, j); F(R, 5, 1, j) = FP(R, 5, nx, j); F(R, 8, 1, j) = FP(R, 8, nx, j);
                F(R, 3, nx, j) = FP(R, 3, 1, j); F(R, 6, nx, j) = FP(R, 6, 1, j); F(R, 7, nx, j) = FP(R, 7, 1, j);
            }
        } else {
        if (R->coords[0] == 0)
            for (int j = 1; j <= ny; ++j) { F(R, 1, 1, j) = FP(R, 3, 1, j); F(R, 5, 1, j) = FP(R, 7, 1, j); F(R, 8, 1, j) = FP(R, 6, 1, j); }
        if (R->coords[0] == w->dims[0] - 1)
            for (int j = 1; j <= ny; ++j) { F(R, 3, nx, j) = FP(R, 1, nx, j); F(R, 6, nx, j) = FP(R, 8, nx, j); F(R, 7, nx, j) = FP(R, 5, nx, j); }
        }
        if (R->coords[1] == 0)
            for (int i = 1; i <= nx; ++i) { F(R, 2, i, 1) = FP(R, 4, i, 1); F(R, 5, i, 1) = FP(R, 7, i, 1); F(R, 6, i, 1) = FP(R, 8, i, 1); }
        if (R->coords[1] == w->dims[1] - 1)
            for (int i = 1; i <= nx; ++i) { F(R, 4, i, ny) = FP(R, 2, i, ny); F(R, 7, i, ny) = FP(R, 5, i, ny); F(R, 8, i, ny) = FP(R, 6, i, ny); }
        if (!w->moving) continue;
        /* moving walls, RB2:790-898: a diagonal population that comes off a wall gets - rho*C/6 with rho of the previous
         * macro(); C = -(ex*U) for a horizontal wall moving at U, -(ey*V) for a vertical wall moving at V, and in the corner
         * cells the population that comes out of the corner itself takes both (RB2:872,880,888,896).  The walls' halves meet
         * at nxHalf / nyHalf of the GLOBAL lattice. */
        const int left = R->coords[0] == 0 && !w->periodic_x, right = R->coords[0] == w->dims[0] - 1 && !w->periodic_x;
        const int bottom = R->coords[1] == 0, top = R->coords[1] == w->dims[1] - 1;
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const int xm = left && i == 1, xp = right && i == nx, ym = bottom && j == 1, yp = top && j == ny;
                if (!(xm | xp | ym | yp)) continue;
                const int gi = R->start[0] + i, gj = R->start[1] + j;
                for (int a = 5; a < Q9; ++a) {
                    const int hx = (ex[a] == 1 && xm) || (ex[a] == -1 && xp);      /* upstream cell beyond a vertical wall   */
                    const int hy = (ey[a] == 1 && ym) || (ey[a] == -1 && yp);      /* upstream cell beyond a horizontal wall */
                    if (!(hx | hy)) continue;
                    static const int opp[Q9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
                    const double cu = -((double)ex[a] * wall_u(w, ey[a] == -1, gi)), cv = -((double)ey[a] * wall_v(w, ex[a] == -1, gj));
                    const double Cw = (hx && hy) ? cu + cv : hy ? cu : cv;
                    F(R, a, i, j) = FP(R, opp[a], i, j) - S(R, rho, i, j) * Cw / 6.0;
                }
            }
    }
}

/* bouncebackT(): evolution_g.F90:79-142.  Adiabatic: g(a) = g_post(opp); constant temperature:
 * g(a) = -g_post(opp) + (4+paraA)/10*Twall  (horizontal walls first, then vertical; D2Q5 has no diagonal, so no overlap) */
void t2_bouncebackT(t2_world *w) {
    const t2_params *p = &w->p;
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny;
        const int on[4] = {R->coords[0] == w->dims[0] - 1, R->coords[0] == 0, R->coords[1] == w->dims[1] - 1, R->coords[1] == 0};
        static const int in_pop[4] = {3, 1, 4, 2}, out_pop[4] = {1, 3, 2, 4};      /* at +x wall: g(3) <- g_post(1) ... */
        static const int order[4] = {3, 2, 1, 0};                                   /* bottom, top, left, right */
        for (int q = 0; q < 4; ++q) {
            const int face = order[q];
            if (!on[face]) continue;
            const int kind = w->bcT[face];
            if (kind == T2_PERIODIC) {      /* VerticalWallsPeriodicalT, acc:1037-1045 */
                for (int j = 1; j <= ny; ++j) G(R, in_pop[face], face == 0 ? nx : 1, j) = GP(R, in_pop[face], face == 0 ? 1 : nx, j);
                continue;
            }
            const double Tw = kind == T2_CONST_HOT ? p->Thot : p->Tcold;
            const int n = face < 2 ? ny : nx;
            for (int t = 1; t <= n; ++t) {
                const int i = face == 0 ? nx : face == 1 ? 1 : t, j = face == 2 ? ny : face == 3 ? 1 : t;
                if (kind == T2_ADIABATIC) G(R, in_pop[face], i, j) = GP(R, out_pop[face], i, j);
                else G(R, in_pop[face], i, j) = -GP(R, out_pop[face], i, j) + (4.0 + p->paraA) / 10.0 * Tw;
            }
        }
        if (w->cornersT && !w->periodic_x)      /* RB2:1086-1106: in a corner cell the population off the vertical wall takes the plate's rule */
            for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                    const int fx = cx ? 0 : 1, fy = cy ? 2 : 3;                  /* faces of this corner */
                    if (!on[fx] || !on[fy]) continue;
                    const int kind = w->bcT[fy];
                    if (kind != T2_CONST_HOT && kind != T2_CONST_COLD) continue;
                    const double Tw = kind == T2_CONST_HOT ? p->Thot : p->Tcold;
                    const int i = cx ? nx : 1, j = cy ? ny : 1;
                    G(R, in_pop[fx], i, j) = -GP(R, out_pop[fx], i, j) + (4.0 + p->paraA) / 10.0 * Tw;
                }
    }
}

/* macro(): evolution_f.F90:328-342 ; macroT(): evolution_g.F90:163-176 */
void t2_macro(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                const double *f = &F(R, 0, i, j);
                double rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
                S(R, rho, i, j) = rho;
                S(R, u, i, j) = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8] + 0.5 * S(R, Fx, i, j)) / rho;
                S(R, v, i, j) = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8] + 0.5 * S(R, Fy, i, j)) / rho;
            }
    }
}
void t2_macroT(t2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                const double *g = &G(R, 0, i, j);
                S(R, T, i, j) = g[0] + g[1] + g[2] + g[3] + g[4];
            }
    }
}

/* check(): check.F90:10-39 (rank sums, then 4 Allreduce in rank order) */
void t2_check(t2_world *w, double *errorU, double *errorT) {
    double t1 = 0.0, t2 = 0.0, t5 = 0.0, t6 = 0.0;
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        double e1 = 0.0, e2 = 0.0, e5 = 0.0, e6 = 0.0;
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                e1 = e1 + (S(R, u, i, j) - S(R, up, i, j)) * (S(R, u, i, j) - S(R, up, i, j)) + (S(R, v, i, j) - S(R, vp, i, j)) * (S(R, v, i, j) - S(R, vp, i, j));
                e2 = e2 + S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j);
                e5 = e5 + fabs(S(R, T, i, j) - S(R, Tp, i, j));
                e6 = e6 + fabs(S(R, T, i, j));
                S(R, up, i, j) = S(R, u, i, j); S(R, vp, i, j) = S(R, v, i, j); S(R, Tp, i, j) = S(R, T, i, j);
            }
        t1 += e1; t2 += e2; t5 += e5; t6 += e6;
    }
    w->errorU = sqrt(t1) / sqrt(t2);
    w->errorT = t5 / t6;
    if (errorU) *errorU = w->errorU;
    if (errorT) *errorT = w->errorT;
}

/* calNuRe()'s three volume sums (NuRe.F90:27-78; the subroutine is written for one rank: i, j are global there):
 * out[0] = sum (i-nxHalf)*v - (j-nyHalf)*u, out[1] = sum v*T, out[2] = sum u*u+v*v, rank by rank in loop order */
void t2_nure_sums(t2_world *w, double out[3]) {
    const int nxHalf = (w->total[0] - 1) / 2 + 1, nyHalf = (w->total[1] - 1) / 2 + 1;     /* module.F90:57 */
    out[0] = out[1] = out[2] = 0.0;
    for (int r = 0; r < w->np; ++r) {
        t2_rank *R = &w->r[r];
        double a = 0.0, b = 0.0, c = 0.0;
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                a = a + (double)(R->start[0] + i - nxHalf) * S(R, v, i, j) - (double)(R->start[1] + j - nyHalf) * S(R, u, i, j);
                b = b + S(R, v, i, j) * S(R, T, i, j);
                c = c + (S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j));
            }
        out[0] += a; out[1] += b; out[2] += c;
    }
}

/* n iterations of the driver loop body: main.F90:84-108 */
void t2_step(t2_world *w, int n) {
    for (int s = 0; s < n; ++s) {
        w->itc += 1;
        t2_collision(w);
        t2_exchange_f(w);
        t2_streaming(w);
        t2_bounceback(w);
        t2_collisionT(w);
        t2_exchange_g(w);
        t2_streamingT(w);
        t2_bouncebackT(w);
        t2_macro(w);
        t2_macroT(w);
    }
}
