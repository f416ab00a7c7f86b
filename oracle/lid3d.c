/*
 * oracle/lid3d.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's D3Q19 MRT
 * lid-driven-cavity hot path (cheryli/MGLC, MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/,
 * "L3" below).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this; the product (mglc_b200/, libmglc.so) never does.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path and its Fortran+MPI
 * sources cannot be compiled in this image (no gfortran / mpif90), so the restatement is pinned to the
 * reference's own SOURCE TEXT: tests/golden/fortran_eval.py machine-evaluates the Fortran statements where
 * they lie under /root/reference and the generators commit the numbers --
 *   per cell     collision.f90:20-189, macro.f90:13-22, initial.f90:66-70  (make_golden_fortran.py ->
 *                ref_fortran_kernels.npz)
 *   whole array  streaming.f90:8-20, bounce_back.f90:6-83 for every wall combination a block can own (the
 *                later wall winning on edges, the lid term with the previous rho), check.f90:9-19
 *                (make_golden_lid3d_fields.py -> ref_fortran_lid3d_fields.npz)
 *   whole run    the SEQUENTIAL program 3d/seq/lid_driven_cavity_3d.f90 evaluated as a whole on 6 x 5 x 4: initial()
 *                and its loop for 1, 2, 12, 14 iterations with check() (make_golden_lid3d_seq_run.py ->
 *                ref_fortran_lid3d_seq_run.npz); this file reproduces f, f_post, rho, u, v, w on 1..8 emulated ranks
 * -- and tests/test_oracle_lid.py requires this file to reproduce them bit for bit.  On top: analytic known
 * answers (M^-1 M = I, M feq = meq, rest equilibrium fixed point, delta-population transport, mass
 * conservation, no NaN leaking from poisoned wall halos) and the reference's implicit seq == MPI contract
 * (decomposition invariance, bit for bit).  The exchange (ex_sendrecv.f90) is MPI calls and cannot be
 * evaluated; it is checked by construction tests and, across real processes, by tests/test_gloo_multiprocess.py.
 *
 * Layout is the reference's: Fortran column-major, AoS with the population index fastest,
 *   f     (0:18, 1:nx,   1:ny,   1:nz  )      L3/initial.f90:43
 *   f_post(0:18, 0:nx+1, 0:ny+1, 0:nz+1)      L3/initial.f90:44
 *   rho,u,v,w,up,vp,wp (1:nx,1:ny,1:nz)       L3/initial.f90:35-41
 * Expressions keep the reference's left-to-right order, divisions stay divisions; build with
 * -ffp-contract=off and without -ffast-math so every operation is one IEEE fp64 rounding.
 * One process emulates all P MPI ranks (same decompose_1d, nbr_surface, nbr_line tables).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19

/* Thread count of every OpenMP loop in this library, set and read back by bench.py's CPU legs: a launcher such as
 * torch.distributed.run exports OMP_NUM_THREADS=1, and the environment is only read when the OpenMP runtime starts. */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* D3Q19 velocity set, L3/commondata.f90:32-40 */
static const int ex[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int ey[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int ez[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};

typedef struct orc_rank {
    int nx, ny, nz;
    int coords[3];
    int start[3];          /* global 0-based offset of local cell 1 */
    int nbr_surface[7];    /* 1..6 : +x -x +y -y +z -z, -1 = MPI_PROC_NULL   L3/main.f90:43-45 */
    int nbr_line[Q];       /* 7..18: rank at coords + e_alpha, -1 = none     L3/main.f90:158-169 */
    double *f, *f_post, *rho, *u, *v, *w, *up, *vp, *wp;
} orc_rank;

typedef struct orc_world {
    int total[3];
    int dims[3];
    int np;
    double Re, rho0, U0, tauf, Snu, Sq;
    int bgk;               /* 1 = the single-relaxation-time alternative, L3/collision.f90:191-198 (commented out there) */
    int itc;
    double errorU;
    orc_rank *r;
} orc_world;

#define F(R, a, i, j, k) ((R)->f[(a) + Q * ((size_t)((i)-1) + (size_t)(R)->nx * ((size_t)((j)-1) + (size_t)(R)->ny * (size_t)((k)-1)))])
#define FP(R, a, i, j, k) ((R)->f_post[(a) + Q * ((size_t)(i) + (size_t)((R)->nx + 2) * ((size_t)(j) + (size_t)((R)->ny + 2) * (size_t)(k)))])
#define S(R, A, i, j, k) ((R)->A[(size_t)((i)-1) + (size_t)(R)->nx * ((size_t)((j)-1) + (size_t)(R)->ny * (size_t)((k)-1))])

/* ---- decomposition ------------------------------------------------------------------------- */

/* MPI_Dims_create(np, 3, dims) with dims = 0 (L3/main.f90:24): balanced factorisation in
 * non-increasing order (2 -> 2,1,1; 4 -> 2,2,1; 8 -> 2,2,2; 6 -> 3,2,1; 12 -> 3,2,2). */
void orc_dims_create(int np, int dims[3]) {
    int best[3] = {np, 1, 1};
    for (int a = 1; a <= np; ++a) {
        if (np % a) continue;
        for (int b = 1; b <= a; ++b) {
            if ((np / a) % b) continue;
            int c = np / a / b;
            if (c > b) continue;
            if (a < best[0] || (a == best[0] && b < best[1])) { best[0] = a; best[1] = b; best[2] = c; }
        }
    }
    dims[0] = best[0]; dims[1] = best[1]; dims[2] = best[2];
}

/* decompose_1d, L3/main.f90:144-155; start = sum of the lower ranks' sizes. */
void orc_decompose_1d(int total_n, int rank, int np, int *local_n, int *start) {
    int n = total_n / np, m = total_n % np;
    *local_n = n + (rank < m ? 1 : 0);
    *start = rank * n + (rank < m ? rank : m);
}

/* MPI_Cart_rank for a row-major Cartesian communicator: rank = (c0*d1 + c1)*d2 + c2 */
static int cart_rank(const int dims[3], const int c[3]) {
    for (int d = 0; d < 3; ++d) if (c[d] < 0 || c[d] >= dims[d]) return -1;
    return (c[0] * dims[1] + c[1]) * dims[2] + c[2];
}

orc_world *orc_world_create(int tnx, int tny, int tnz, int np, const int *dims_or_null,
                            double Re, double U0, double rho0) {
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    w->total[0] = tnx; w->total[1] = tny; w->total[2] = tnz;
    w->np = np;
    if (dims_or_null && dims_or_null[0] > 0) memcpy(w->dims, dims_or_null, 3 * sizeof(int));
    else orc_dims_create(np, w->dims);
    w->Re = Re; w->rho0 = rho0; w->U0 = U0;
    /* L3/commondata.f90:9,42 */
    w->tauf = U0 * (double)tnx / Re * 3.0 + 0.5;
    w->Snu = 1.0 / w->tauf;
    w->Sq = 8.0 * (2.0 * w->tauf - 1.0) / (8.0 * w->tauf - 1.0);
    w->r = (orc_rank *)calloc((size_t)np, sizeof(orc_rank));
    for (int c0 = 0; c0 < w->dims[0]; ++c0)
    for (int c1 = 0; c1 < w->dims[1]; ++c1)
    for (int c2 = 0; c2 < w->dims[2]; ++c2) {
        int c[3] = {c0, c1, c2};
        orc_rank *R = &w->r[cart_rank(w->dims, c)];
        memcpy(R->coords, c, sizeof c);
        orc_decompose_1d(tnx, c0, w->dims[0], &R->nx, &R->start[0]);
        orc_decompose_1d(tny, c1, w->dims[1], &R->ny, &R->start[1]);
        orc_decompose_1d(tnz, c2, w->dims[2], &R->nz, &R->start[2]);
        /* MPI_Cart_shift(dir, +1, source, dest): nbr_surface(2)=source(-), (1)=dest(+)  L3/main.f90:43-45 */
        for (int d = 0; d < 3; ++d) {
            int p[3] = {c0, c1, c2}, m[3] = {c0, c1, c2};
            p[d] += 1; m[d] -= 1;
            R->nbr_surface[2 * d + 1] = cart_rank(w->dims, p);
            R->nbr_surface[2 * d + 2] = cart_rank(w->dims, m);
        }
        for (int a = 7; a < Q; ++a) {
            int n[3] = {c0 + ex[a], c1 + ey[a], c2 + ez[a]};
            R->nbr_line[a] = cart_rank(w->dims, n);
        }
        size_t n = (size_t)R->nx * R->ny * R->nz;
        size_t nh = (size_t)(R->nx + 2) * (R->ny + 2) * (R->nz + 2);
        R->f = (double *)malloc(Q * n * sizeof(double));
        R->f_post = (double *)malloc(Q * nh * sizeof(double));
        R->rho = (double *)malloc(n * sizeof(double));
        R->u = (double *)malloc(n * sizeof(double));
        R->v = (double *)malloc(n * sizeof(double));
        R->w = (double *)malloc(n * sizeof(double));
        R->up = (double *)malloc(n * sizeof(double));
        R->vp = (double *)malloc(n * sizeof(double));
        R->wp = (double *)malloc(n * sizeof(double));
        /* The reference leaves f_post uninitialised (L3/initial.f90:44).  Poison it with NaN so a
         * test can prove that no halo value at a physical wall ever survives into f. */
        for (size_t q = 0; q < Q * nh; ++q) R->f_post[q] = NAN;
    }
    return w;
}

void orc_world_destroy(orc_world *w) {
    if (!w) return;
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
        free(R->f); free(R->f_post); free(R->rho); free(R->u); free(R->v); free(R->w);
        free(R->up); free(R->vp); free(R->wp);
    }
    free(w->r); free(w);
}

/* which: 0 f, 1 f_post, 2 rho, 3 u, 4 v, 5 w, 6 up, 7 vp, 8 wp */
double *orc_rank_ptr(orc_world *w, int r, int which) {
    orc_rank *R = &w->r[r];
    switch (which) {
        case 0: return R->f; case 1: return R->f_post; case 2: return R->rho; case 3: return R->u;
        case 4: return R->v; case 5: return R->w; case 6: return R->up; case 7: return R->vp;
        case 8: return R->wp; default: return NULL;
    }
}

/* out[0..2]=n, [3..5]=coords, [6..8]=start, [9..14]=nbr_surface(1..6), [15..26]=nbr_line(7..18) */
void orc_rank_info(orc_world *w, int r, int *out) {
    orc_rank *R = &w->r[r];
    out[0] = R->nx; out[1] = R->ny; out[2] = R->nz;
    for (int d = 0; d < 3; ++d) { out[3 + d] = R->coords[d]; out[6 + d] = R->start[d]; }
    for (int s = 1; s <= 6; ++s) out[8 + s] = R->nbr_surface[s];
    for (int a = 7; a < Q; ++a) out[8 + a] = R->nbr_line[a];
}

void orc_world_info(orc_world *w, int *dims, double *params /* tauf,Snu,Sq,errorU */, int *itc) {
    memcpy(dims, w->dims, 3 * sizeof(int));
    params[0] = w->tauf; params[1] = w->Snu; params[2] = w->Sq; params[3] = w->errorU;
    *itc = w->itc;
}

/* weights, L3/commondata.f90:29-31 */
static const double omega[Q] = {1.0 / 3.0,
    1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0,
    1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0,
    1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};

/* ---- initial(), L3/initial.f90:1-76 ---------------------------------------------------------- */
void orc_initial(orc_world *w) {
    w->itc = 0;
    w->errorU = 100.0;
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
        size_t n = (size_t)R->nx * R->ny * R->nz;
        for (size_t q = 0; q < n; ++q) {
            R->rho[q] = w->rho0; R->u[q] = 0.0; R->v[q] = 0.0; R->w[q] = 0.0;
            R->up[q] = 0.0; R->vp[q] = 0.0; R->wp[q] = 0.0;
        }
        if (R->coords[2] == w->dims[2] - 1)        /* top boundary, :55-61 */
            for (int j = 1; j <= R->ny; ++j)
                for (int i = 1; i <= R->nx; ++i) S(R, u, i, j, R->nz) = w->U0;
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i) {
            double u = S(R, u, i, j, k), v = S(R, v, i, j, k), ww = S(R, w, i, j, k), rho = S(R, rho, i, j, k);
            double us2 = u * u + v * v + ww * ww;
            for (int a = 0; a < Q; ++a) {
                double un = u * (double)ex[a] + v * (double)ey[a] + ww * (double)ez[a];
                F(R, a, i, j, k) = rho * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
            }
        }
    }
}

/* ---- collision(), L3/collision.f90:1-205 ----------------------------------------------------- */
/* m = M f, hand-expanded forward transform */
void orc_moments(const double *f, double *m) {
    /* forward transform, :20-70 */
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12]
         + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[1] = -30.0 * f[0] - 11.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + 8.0 * (f[7] + f[8] + f[9] + f[10] + f[11] + f[12])
         + 8.0 * (f[13] + f[14] + f[15] + f[16] + f[17] + f[18]);
    m[2] = 12.0 * f[0] - 4.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[3] = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[4] = -4.0 * (f[1] - f[2]) + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[5] = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[6] = -4.0 * (f[3] - f[4]) + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[7] = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[8] = -4.0 * (f[5] - f[6]) + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[9] = 2.0 * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12]
         + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[10] = -4.0 * (f[1] + f[2]) + 2.0 * (f[3] + f[4] + f[5] + f[6]) + f[7] + f[8] + f[9] + f[10] + f[11]
          + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[11] = f[3] + f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];
    m[12] = -2.0 * (f[3] + f[4] - f[5] - f[6]) + f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];
    m[13] = f[7] - f[8] - f[9] + f[10];
    m[14] = f[15] - f[16] - f[17] + f[18];
    m[15] = f[11] - f[12] - f[13] + f[14];
    m[16] = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    m[17] = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    m[18] = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];
}

void orc_meq(double rho, double u, double v, double w, double *meq) {
    /* equilibrium moments, :73-91.  meq(12) carries NO rho factor (:85) -- reference quirk, kept. */
    meq[0] = rho;
    meq[1] = rho * (-11.0 + 19.0 * (u * u + v * v + w * w));
    meq[2] = rho * (3.0 - 11.0 / 2.0 * (u * u + v * v + w * w));
    meq[3] = rho * u;
    meq[4] = -2.0 / 3.0 * rho * u;
    meq[5] = rho * v;
    meq[6] = -2.0 / 3.0 * rho * v;
    meq[7] = rho * w;
    meq[8] = -2.0 / 3.0 * rho * w;
    meq[9] = rho * (2.0 * u * u - v * v - w * w);
    meq[10] = -1.0 / 2.0 * rho * (2.0 * u * u - v * v - w * w);
    meq[11] = rho * (v * v - w * w);
    meq[12] = -1.0 / 2.0 * (v * v - w * w);
    meq[13] = rho * u * v;
    meq[14] = rho * v * w;
    meq[15] = rho * u * w;
    meq[16] = 0.0; meq[17] = 0.0; meq[18] = 0.0;
}

/* f = M^-1 m, hand-expanded inverse transform (fp = output populations, mp = input moments) */
void orc_inverse(const double *mp, double *fp) {
    /* inverse transform, :118-189 */
    fp[0] = mp[0] / 19.0 - 5.0 / 399.0 * mp[1] + mp[2] / 21.0;
    fp[1] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + mp[3] / 10.0 - mp[4] / 10.0 + mp[9] / 18.0 - mp[10] / 18.0;
    fp[2] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - mp[3] / 10.0 + mp[4] / 10.0 + mp[9] / 18.0 - mp[10] / 18.0;
    fp[3] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + mp[5] / 10.0 - mp[6] / 10.0 - mp[9] / 36.0 + mp[10] / 36.0 + mp[11] / 12.0 - mp[12] / 12.0;
    fp[4] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - mp[5] / 10.0 + mp[6] / 10.0 - mp[9] / 36.0 + mp[10] / 36.0 + mp[11] / 12.0 - mp[12] / 12.0;
    fp[5] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + mp[7] / 10.0 - mp[8] / 10.0 - mp[9] / 36.0 + mp[10] / 36.0 - mp[11] / 12.0 + mp[12] / 12.0;
    fp[6] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - mp[7] / 10.0 + mp[8] / 10.0 - mp[9] / 36.0 + mp[10] / 36.0 - mp[11] / 12.0 + mp[12] / 12.0;
    fp[7] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[3] / 10.0 + mp[4] / 40.0 + mp[5] / 10.0 + mp[6] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 + mp[13] / 4.0 + mp[16] / 8.0 - mp[17] / 8.0;
    fp[8] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[3] / 10.0 - mp[4] / 40.0 + mp[5] / 10.0 + mp[6] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 - mp[13] / 4.0 - mp[16] / 8.0 - mp[17] / 8.0;
    fp[9] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[3] / 10.0 + mp[4] / 40.0 - mp[5] / 10.0 - mp[6] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 - mp[13] / 4.0 + mp[16] / 8.0 + mp[17] / 8.0;
    fp[10] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[3] / 10.0 - mp[4] / 40.0 - mp[5] / 10.0 - mp[6] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 + mp[13] / 4.0 - mp[16] / 8.0 + mp[17] / 8.0;
    fp[11] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[3] / 10.0 + mp[4] / 40.0 + mp[7] / 10.0 + mp[8] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 + mp[15] / 4.0 - mp[16] / 8.0 + mp[18] / 8.0;
    fp[12] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[3] / 10.0 - mp[4] / 40.0 + mp[7] / 10.0 + mp[8] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 - mp[15] / 4.0 + mp[16] / 8.0 + mp[18] / 8.0;
    fp[13] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[3] / 10.0 + mp[4] / 40.0 - mp[7] / 10.0 - mp[8] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 - mp[15] / 4.0 - mp[16] / 8.0 - mp[18] / 8.0;
    fp[14] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[3] / 10.0 - mp[4] / 40.0 - mp[7] / 10.0 - mp[8] / 40.0 + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 + mp[15] / 4.0 + mp[16] / 8.0 - mp[18] / 8.0;
    fp[15] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[5] / 10.0 + mp[6] / 40.0 + mp[7] / 10.0 + mp[8] / 40.0 - mp[9] / 18.0 - mp[10] / 36.0 + mp[14] / 4.0 + mp[17] / 8.0 - mp[18] / 8.0;
    fp[16] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[5] / 10.0 - mp[6] / 40.0 + mp[7] / 10.0 + mp[8] / 40.0 - mp[9] / 18.0 - mp[10] / 36.0 - mp[14] / 4.0 - mp[17] / 8.0 - mp[18] / 8.0;
    fp[17] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + mp[5] / 10.0 + mp[6] / 40.0 - mp[7] / 10.0 - mp[8] / 40.0 - mp[9] / 18.0 - mp[10] / 36.0 - mp[14] / 4.0 + mp[17] / 8.0 + mp[18] / 8.0;
    fp[18] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - mp[5] / 10.0 - mp[6] / 40.0 - mp[7] / 10.0 - mp[8] / 40.0 - mp[9] / 18.0 - mp[10] / 36.0 + mp[14] / 4.0 - mp[17] / 8.0 + mp[18] / 8.0;
}

/* One cell: f[19], rho,u,v,w -> fp[19]. */
void orc_collide_cell(const double *f, double rho, double u, double v, double w,
                      double Snu, double Sq, double *fp) {
    double m[Q], meq[Q], s[Q], mp[Q];
    orc_moments(f, m);
    orc_meq(rho, u, v, w, meq);
    /* relaxation rates, :94-112 */
    s[0] = 0.0; s[1] = Snu; s[2] = Snu; s[3] = 0.0; s[4] = Sq; s[5] = 0.0; s[6] = Sq; s[7] = 0.0;
    s[8] = Sq; s[9] = Snu; s[10] = Snu; s[11] = Snu; s[12] = Snu; s[13] = Snu; s[14] = Snu;
    s[15] = Snu; s[16] = Sq; s[17] = Sq; s[18] = Sq;
    for (int a = 0; a < Q; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);   /* :114-116 */
    orc_inverse(mp, fp);
}

/* The BGK alternative the reference keeps as a comment block at the end of the cell loop
 * (L3/collision.f90:191-198): feq as in initial(), one rate Snu for every population. */
void orc_collide_cell_bgk(const double *f, double rho, double u, double v, double w, double Snu, double *fp) {
    double us2 = u * u + v * v + w * w;
    for (int a = 0; a < Q; ++a) {
        double un = u * (double)ex[a] + v * (double)ey[a] + w * (double)ez[a];
        double feq = rho * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
        fp[a] = f[a] - Snu * (f[a] - feq);
    }
}

void orc_world_set_bgk(orc_world *w, int on) { w->bgk = on ? 1 : 0; }

void orc_collision(orc_world *w) {
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
        if (w->bgk) {
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= R->nz; ++k)
            for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                orc_collide_cell_bgk(&F(R, 0, i, j, k), S(R, rho, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k),
                                     S(R, w, i, j, k), w->Snu, &FP(R, 0, i, j, k));
            continue;
        }
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i)
            orc_collide_cell(&F(R, 0, i, j, k), S(R, rho, i, j, k), S(R, u, i, j, k), S(R, v, i, j, k),
                             S(R, w, i, j, k), w->Snu, w->Sq, &FP(R, 0, i, j, k));
    }
}

/* ---- message_passing_sendrecv(), L3/ex_sendrecv.f90:1-129 ------------------------------------ */
/* A Sendrecv pair (send to `dest`, receive from `source`) across all ranks is the same as: every
 * rank with a valid `dest` copies its send region into dest's receive region. */
static void copy_face(orc_world *w, int dir /*1..6*/, const int pops[5]) {
    for (int r = 0; r < w->np; ++r) {
        orc_rank *Sx = &w->r[r];
        int d = Sx->nbr_surface[dir];
        if (d < 0) continue;
        orc_rank *D = &w->r[d];
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

/* edge for diagonal population a: send the line at the (e_a)-most corner row of the interior to
 * nbr_line(a), where it lands in the opposite halo row.  :64-123 */
static void copy_edge(orc_world *w, int a) {
    for (int r = 0; r < w->np; ++r) {
        orc_rank *Sx = &w->r[r];
        int d = Sx->nbr_line[a];
        if (d < 0) continue;
        orc_rank *D = &w->r[d];
        /* source index per dim: e=+1 -> n, e=-1 -> 1, e=0 -> runs 1..n ; dest: e=+1 -> 0, e=-1 -> n+1 */
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

void orc_exchange(orc_world *w) {
    static const int px[5] = {1, 7, 9, 11, 13}, mx[5] = {2, 8, 10, 12, 14};
    static const int py[5] = {3, 7, 8, 15, 17}, my[5] = {4, 9, 10, 16, 18};
    static const int pz[5] = {5, 11, 12, 15, 16}, mz[5] = {6, 13, 14, 17, 18};
    copy_face(w, 1, px); copy_face(w, 2, mx);
    copy_face(w, 3, py); copy_face(w, 4, my);
    copy_face(w, 5, pz); copy_face(w, 6, mz);
    static const int order[12] = {7, 10, 9, 8, 11, 14, 13, 12, 15, 18, 17, 16};
    for (int q = 0; q < 12; ++q) copy_edge(w, order[q]);
}

/* ---- streaming(), L3/streaming.f90:1-23 ------------------------------------------------------ */
void orc_streaming(orc_world *w) {
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i)
            for (int a = 0; a < Q; ++a)
                F(R, a, i, j, k) = FP(R, a, i - ex[a], j - ey[a], k - ez[a]);
    }
}

/* ---- bounceback(), L3/bounce_back.f90:1-86 --------------------------------------------------- */
void orc_bounceback(orc_world *w) {
    const double U0 = w->U0;
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny, nz = R->nz;
        if (R->coords[0] == 0)
            for (int k = 1; k <= nz; ++k) for (int j = 1; j <= ny; ++j) {
                F(R, 1, 1, j, k) = FP(R, 2, 1, j, k);   F(R, 7, 1, j, k) = FP(R, 10, 1, j, k);
                F(R, 9, 1, j, k) = FP(R, 8, 1, j, k);   F(R, 11, 1, j, k) = FP(R, 14, 1, j, k);
                F(R, 13, 1, j, k) = FP(R, 12, 1, j, k);
            }
        if (R->coords[0] == w->dims[0] - 1)
            for (int k = 1; k <= nz; ++k) for (int j = 1; j <= ny; ++j) {
                F(R, 2, nx, j, k) = FP(R, 1, nx, j, k);  F(R, 10, nx, j, k) = FP(R, 7, nx, j, k);
                F(R, 8, nx, j, k) = FP(R, 9, nx, j, k);  F(R, 14, nx, j, k) = FP(R, 11, nx, j, k);
                F(R, 12, nx, j, k) = FP(R, 13, nx, j, k);
            }
        if (R->coords[1] == 0)
            for (int k = 1; k <= nz; ++k) for (int i = 1; i <= nx; ++i) {
                F(R, 3, i, 1, k) = FP(R, 4, i, 1, k);   F(R, 7, i, 1, k) = FP(R, 10, i, 1, k);
                F(R, 8, i, 1, k) = FP(R, 9, i, 1, k);   F(R, 15, i, 1, k) = FP(R, 18, i, 1, k);
                F(R, 17, i, 1, k) = FP(R, 16, i, 1, k);
            }
        if (R->coords[1] == w->dims[1] - 1)
            for (int k = 1; k <= nz; ++k) for (int i = 1; i <= nx; ++i) {
                F(R, 4, i, ny, k) = FP(R, 3, i, ny, k);  F(R, 10, i, ny, k) = FP(R, 7, i, ny, k);
                F(R, 9, i, ny, k) = FP(R, 8, i, ny, k);  F(R, 18, i, ny, k) = FP(R, 15, i, ny, k);
                F(R, 16, i, ny, k) = FP(R, 17, i, ny, k);
            }
        if (R->coords[2] == 0)
            for (int j = 1; j <= ny; ++j) for (int i = 1; i <= nx; ++i) {
                F(R, 5, i, j, 1) = FP(R, 6, i, j, 1);   F(R, 11, i, j, 1) = FP(R, 14, i, j, 1);
                F(R, 12, i, j, 1) = FP(R, 13, i, j, 1); F(R, 15, i, j, 1) = FP(R, 18, i, j, 1);
                F(R, 16, i, j, 1) = FP(R, 17, i, j, 1);
            }
        if (R->coords[2] == w->dims[2] - 1)       /* moving lid, :72-83 */
            for (int j = 1; j <= ny; ++j) for (int i = 1; i <= nx; ++i) {
                F(R, 6, i, j, nz) = FP(R, 5, i, j, nz);
                F(R, 14, i, j, nz) = FP(R, 11, i, j, nz) - S(R, rho, i, j, nz) / 6.0 * (U0);
                F(R, 13, i, j, nz) = FP(R, 12, i, j, nz) - S(R, rho, i, j, nz) / 6.0 * (-U0);
                F(R, 18, i, j, nz) = FP(R, 15, i, j, nz);
                F(R, 17, i, j, nz) = FP(R, 16, i, j, nz);
            }
    }
}

/* ---- macro(), L3/macro.f90:1-28 -------------------------------------------------------------- */
void orc_macro(orc_world *w) {
    for (int r = 0; r < w->np; ++r) {
        orc_rank *R = &w->r[r];
#pragma omp parallel for schedule(static)
        for (int k = 1; k <= R->nz; ++k)
        for (int j = 1; j <= R->ny; ++j)
        for (int i = 1; i <= R->nx; ++i) {
            double rho = 0.0, u = 0.0, v = 0.0, ww = 0.0;
            for (int a = 0; a < Q; ++a) {
                double fa = F(R, a, i, j, k);
                rho = rho + fa;
                u = u + fa * (double)ex[a];
                v = v + fa * (double)ey[a];
                ww = ww + fa * (double)ez[a];
            }
            S(R, rho, i, j, k) = rho;
            S(R, u, i, j, k) = u / rho;
            S(R, v, i, j, k) = v / rho;
            S(R, w, i, j, k) = ww / rho;
        }
    }
}

/* ---- check(), L3/check.f90:1-37.  error1 has no w term (:15) -- reference quirk, kept. -------- */
double orc_check(orc_world *w) {
    double total1 = 0.0, total2 = 0.0;
    for (int r = 0; r < w->np; ++r) {       /* Allreduce(SUM) modelled as a rank-ordered sum */
        orc_rank *R = &w->r[r];
        double e1 = 0.0, e2 = 0.0;
        size_t n = (size_t)R->nx * R->ny * R->nz;
        for (size_t q = 0; q < n; ++q) {
            double u = R->u[q], v = R->v[q], ww = R->w[q];
            e1 = e1 + (u - R->up[q]) * (u - R->up[q]) + (v - R->vp[q]) * (v - R->vp[q]);
            e2 = e2 + u * u + v * v + ww * ww;
        }
        memcpy(R->up, R->u, n * sizeof(double));
        memcpy(R->vp, R->v, n * sizeof(double));
        memcpy(R->wp, R->w, n * sizeof(double));
        total1 += e1; total2 += e2;
    }
    w->errorU = sqrt(total1) / sqrt(total2);
    return w->errorU;
}

/* n iterations of the driver loop body, L3/main.f90:85-103 (without the convergence exit) */
void orc_step(orc_world *w, int n) {
    for (int s = 0; s < n; ++s) {
        w->itc += 1;
        orc_collision(w);
        orc_exchange(w);
        orc_streaming(w);
        orc_bounceback(w);
        orc_macro(w);
    }
}
