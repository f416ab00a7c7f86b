/*
 * oracle/lid2d.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's 2-D D2Q9 MRT lid-driven
 * cavity, in its three shipped forms:
 *   L2C  MPI/Lid_driven_cavity/c/lid_driven_cavity.c                      (plain C, one domain, 200 x 200)
 *   L2F  MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/ (.f90)   (Fortran + MPI, 2-D Cartesian blocks, 201 x 201)
 *   L2I  MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90  (sequential, incompressible model, 257 x 257)
 * Only tests/ and __graft_entry__.smoke() may load this; the product never does.
 *
 * PARITY PIN: L2C is the one LBM program of the reference this image can compile.  oracle/Makefile `ref` builds it
 * unmodified into oracle/_ref/liblid2d_ref.so; tests/test_oracle_lid2d.py runs the reference's own initial() /
 * collision() / streaming() / boundary() / macro() / check() and requires this restatement (variant L2_C) to reproduce
 * f, f_post, rho, u, v bit for bit, which also proves the equivalence used below: the reference's push streaming with
 * periodic wrap followed by boundary() (c:262-313) == pull streaming from halo'd f_post followed by bounceback().
 * L2F (variant L2_F) differs from L2C only in the rounding of collision() (grouped sums, per-term divisions,
 * meq(8) = rho*u*v instead of u*v) and in check(); it is pinned through the Fortran-text evaluator
 * (tests/golden/make_golden_lid2d.py: collision, macro, feq per cell; make_golden_lid2d_fields.py: streaming,
 * bounceback for the block positions that change the owned walls, the check() sums as whole arrays) and the
 * seq == MPI contract (P ranks == 1 rank).
 * L2I (variant L2_I) is L2F with the incompressible equilibrium: meq without the rho factors (inc:195-202), u, v = the
 * momentum sums undivided (inc:307-308), the lid term without rho (inc:291-292), initial() with rho = 0 and f = omega*(...)
 * (inc:137,160; so the first collision() relaxes towards meq(1) = 3|u|^2, meq(2) = -3|u|^2), and check() as a ratio of
 * sums of dsqrt (inc:327-335).  Pinned through the same evaluator (make_golden_lid2d_incomp.py: every subroutine as whole
 * arrays).  The program is sequential; on P ranks the restatement exchanges halos as L2F does (P ranks == 1 rank).
 * L2S (variant L2_S) is L2C with its own model switch set to SRT (c:13-14, collision c:160-176: single-relaxation-time BGK,
 * f_post = f - 1.0/tau*(f - feq)); pinned live and through committed outputs to the same program compiled with that switch
 * (oracle/_ref/liblid2d_srt_ref.so; make_golden_lid2d_srt.py).
 *
 * Layout is L2F's: column-major, population index fastest: f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), rho,u,v,up,vp(nx,ny)
 * (initial.f90:30-38); tests transpose when they compare with the C program's f[NX][NY][9], rho[NX][NY].
 * Left-to-right evaluation, true divisions, -ffp-contract=off: every operation is one IEEE fp64 rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q9 9
enum { L2_C = 0, L2_F = 1, L2_I = 2, L2_S = 3 };
#define IS_C(v) ((v) == L2_C || (v) == L2_S)

/* commondata.f90:25-27 == c:24-25 */
static const int ex[Q9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int ey[Q9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

typedef struct l2_rank {
    int nx, ny, coords[2], start[2];
    int nbr[4];          /* right(+x), left(-x), top(+y), bottom(-y); -1 = MPI_PROC_NULL   main.f90:46-47 */
    int cnr[4];          /* top_right(5), top_left(6), bottom_left(7), bottom_right(8)     MPI_Cart_find_corners */
    double *f, *f_post, *rho, *u, *v, *up, *vp;
} l2_rank;

typedef struct l2_world {
    int total[2], dims[2], np, variant, itc;
    double Re, U0, rho0, tauf, Snu, Sq, errorU;
    l2_rank *r;
} l2_world;

#define F(R, a, i, j) ((R)->f[(a) + Q9 * ((size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1))])
#define FP(R, a, i, j) ((R)->f_post[(a) + Q9 * ((size_t)(i) + (size_t)((R)->nx + 2) * (size_t)(j))])
#define S(R, A, i, j) ((R)->A[(size_t)((i)-1) + (size_t)(R)->nx * (size_t)((j)-1)])

/* MPI_Dims_create(np, 2, dims) with dims = 0 (main.f90:28): balanced, non-increasing */
void l2_dims_create(int np, int dims[2]) {
    int best = np;
    for (int a = 1; a <= np; ++a)
        if (np % a == 0 && a >= np / a && a < best) best = a;
    dims[0] = best; dims[1] = np / best;
}
static void decompose_1d(int total_n, int rank, int np, int *local_n, int *start) {   /* main.f90:173-185 */
    int n = total_n / np, m = total_n % np;
    *local_n = n + (rank < m ? 1 : 0);
    *start = rank * n + (rank < m ? rank : m);
}
static int cart_rank(const int dims[2], int c0, int c1) {
    if (c0 < 0 || c0 >= dims[0] || c1 < 0 || c1 >= dims[1]) return -1;
    return c0 * dims[1] + c1;
}

l2_world *l2_world_create(int tnx, int tny, int np, const int *dims_or_null, int variant, double Re, double U0, double rho0) {
    l2_world *w = (l2_world *)calloc(1, sizeof(l2_world));
    w->total[0] = tnx; w->total[1] = tny; w->np = np; w->variant = variant;
    if (dims_or_null && dims_or_null[0] > 0) memcpy(w->dims, dims_or_null, 2 * sizeof(int));
    else l2_dims_create(np, w->dims);
    w->Re = Re; w->U0 = U0; w->rho0 = rho0;
    /* commondata.f90:9,31  ==  c:96-101 (nu = u_zero*height/Re; tau = 3*nu+0.5; the products commute) */
    w->tauf = U0 * (double)tnx / Re * 3.0 + 0.5;
    w->Snu = 1.0 / w->tauf;
    w->Sq = 8.0 * (2.0 * w->tauf - 1.0) / (8.0 * w->tauf - 1.0);
    w->r = (l2_rank *)calloc((size_t)np, sizeof(l2_rank));
    for (int c0 = 0; c0 < w->dims[0]; ++c0)
        for (int c1 = 0; c1 < w->dims[1]; ++c1) {
            l2_rank *R = &w->r[cart_rank(w->dims, c0, c1)];
            R->coords[0] = c0; R->coords[1] = c1;
            decompose_1d(tnx, c0, w->dims[0], &R->nx, &R->start[0]);
            decompose_1d(tny, c1, w->dims[1], &R->ny, &R->start[1]);
            R->nbr[0] = cart_rank(w->dims, c0 + 1, c1); R->nbr[1] = cart_rank(w->dims, c0 - 1, c1);
            R->nbr[2] = cart_rank(w->dims, c0, c1 + 1); R->nbr[3] = cart_rank(w->dims, c0, c1 - 1);
            for (int a = 5; a < Q9; ++a) R->cnr[a - 5] = cart_rank(w->dims, c0 + ex[a], c1 + ey[a]);
            size_t n = (size_t)R->nx * R->ny, nh = (size_t)(R->nx + 2) * (R->ny + 2);
            R->f = (double *)calloc(Q9 * n, sizeof(double));
            R->f_post = (double *)calloc(Q9 * nh, sizeof(double));
            R->rho = (double *)calloc(n, sizeof(double)); R->u = (double *)calloc(n, sizeof(double));
            R->v = (double *)calloc(n, sizeof(double)); R->up = (double *)calloc(n, sizeof(double));
            R->vp = (double *)calloc(n, sizeof(double));
        }
    return w;
}
void l2_world_destroy(l2_world *w) {
    if (!w) return;
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        free(R->f); free(R->f_post); free(R->rho); free(R->u); free(R->v); free(R->up); free(R->vp);
    }
    free(w->r); free(w);
}
void l2_world_info(l2_world *w, int dims[2], double par[3]) {
    dims[0] = w->dims[0]; dims[1] = w->dims[1];
    par[0] = w->tauf; par[1] = w->Snu; par[2] = w->Sq;
}
void l2_rank_info(l2_world *w, int r, int info[14]) {
    l2_rank *R = &w->r[r];
    info[0] = R->nx; info[1] = R->ny; info[2] = R->coords[0]; info[3] = R->coords[1]; info[4] = R->start[0]; info[5] = R->start[1];
    memcpy(info + 6, R->nbr, sizeof R->nbr); memcpy(info + 10, R->cnr, sizeof R->cnr);
}
double *l2_rank_ptr(l2_world *w, int r, int which) {
    l2_rank *R = &w->r[r];
    double *p[] = {R->f, R->f_post, R->rho, R->u, R->v, R->up, R->vp};
    return p[which];
}

/* initial(): initial.f90:40-66 == c:123-151 */
void l2_initial(l2_world *w) {
    static const double omega[Q9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
    w->itc = 0; w->errorU = 100.0;
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                S(R, rho, i, j) = w->variant == L2_I ? 0.0 : w->rho0;      /* inc:137 */
                S(R, u, i, j) = 0.0; S(R, v, i, j) = 0.0; S(R, up, i, j) = 0.0; S(R, vp, i, j) = 0.0;
            }
        if (R->coords[1] == w->dims[1] - 1)
            for (int i = 1; i <= R->nx; ++i) S(R, u, i, R->ny) = w->U0;
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                double us2 = S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j);
                for (int a = 0; a < Q9; ++a) {
                    double un = S(R, u, i, j) * (double)ex[a] + S(R, v, i, j) * (double)ey[a];
                    if (w->variant == L2_I) F(R, a, i, j) = omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);      /* inc:160 */
                    else F(R, a, i, j) = S(R, rho, i, j) * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
                }
            }
    }
}

/* collision() of one cell.  variant L2_S: c:160-176 ; variant L2_C: c:186-255 ; variant L2_F: evolution.f90:15-70 ; variant L2_I: inc:183-233 */
void l2_collide_cell(int variant, const double *f, double rho, double u, double v, double Snu, double Sq, double *fp) {
    double m[Q9], meq[Q9], mp[Q9];
    const double s[Q9] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, Snu, Snu};
    if (variant == L2_S) {                               /* c:160-176; 1.0/tau is s_nu (c:99) */
        static const double omega[Q9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
        double u2 = u * u + v * v;
        for (int a = 0; a < Q9; ++a) {
            double ue = u * (double)ex[a] + v * (double)ey[a];
            double feq = rho * omega[a] * (1.0 + 3.0 * ue + 4.5 * ue * ue - 1.5 * u2);
            fp[a] = f[a] - Snu * (f[a] - feq);
        }
        return;
    }
    if (variant == L2_C) {
        m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        m[1] = -4 * f[0] - f[1] - f[2] - f[3] - f[4] + 2 * f[5] + 2 * f[6] + 2 * f[7] + 2 * f[8];
        m[2] = 4 * f[0] - 2 * f[1] - 2 * f[2] - 2 * f[3] - 2 * f[4] + f[5] + f[6] + f[7] + f[8];
        m[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
        m[4] = -2 * f[1] + 2 * f[3] + f[5] - f[6] - f[7] + f[8];
        m[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
        m[6] = -2 * f[2] + 2 * f[4] + f[5] + f[6] - f[7] - f[8];
        m[7] = f[1] - f[2] + f[3] - f[4];
        m[8] = f[5] - f[6] + f[7] - f[8];
        meq[0] = rho;
        meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
        meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
        meq[3] = rho * u;
        meq[4] = -1.0 * rho * u;
        meq[5] = rho * v;
        meq[6] = -1.0 * rho * v;
        meq[7] = rho * (u * u - v * v);
        meq[8] = u * v;                                  /* c:211 -- no rho factor */
        for (int a = 0; a < Q9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);
        fp[0] = (mp[0] - mp[1] + mp[2]) / 9.0;
        fp[1] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] + 6.0 * mp[3] - 6.0 * mp[4] + 9.0 * mp[7]) / 36.0;
        fp[2] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] + 6.0 * mp[5] - 6.0 * mp[6] - 9.0 * mp[7]) / 36.0;
        fp[3] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] - 6.0 * mp[3] + 6.0 * mp[4] + 9.0 * mp[7]) / 36.0;
        fp[4] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] - 6.0 * mp[5] + 6.0 * mp[6] - 9.0 * mp[7]) / 36.0;
        fp[5] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] + 6.0 * mp[3] + 3.0 * mp[4] + 6.0 * mp[5] + 3.0 * mp[6] + 9.0 * mp[8]) / 36.0;
        fp[6] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] - 6.0 * mp[3] - 3.0 * mp[4] + 6.0 * mp[5] + 3.0 * mp[6] - 9.0 * mp[8]) / 36.0;
        fp[7] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] - 6.0 * mp[3] - 3.0 * mp[4] - 6.0 * mp[5] - 3.0 * mp[6] + 9.0 * mp[8]) / 36.0;
        fp[8] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] + 6.0 * mp[3] + 3.0 * mp[4] - 6.0 * mp[5] - 3.0 * mp[6] - 9.0 * mp[8]) / 36.0;
    } else {
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
        if (variant == L2_I) {                           /* inc:194-202 */
            meq[1] = -2.0 * rho + 3.0 * (u * u + v * v);
            meq[2] = rho - 3.0 * (u * u + v * v);
            meq[3] = u;
            meq[4] = -u;
            meq[5] = v;
            meq[6] = -v;
            meq[7] = u * u - v * v;
            meq[8] = u * v;
        } else {
            meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
            meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
            meq[3] = rho * u;
            meq[4] = -rho * u;
            meq[5] = rho * v;
            meq[6] = -rho * v;
            meq[7] = rho * (u * u - v * v);
            meq[8] = rho * (u * v);
        }
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
}
void l2_collision(l2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                l2_collide_cell(w->variant, &F(R, 0, i, j), S(R, rho, i, j), S(R, u, i, j), S(R, v, i, j), w->Snu, w->Sq, &FP(R, 0, i, j));
    }
}

/* message_passing_sendrecv(): ex_sendrecv.f90:9-78 -- 3 populations per face over the interior range, 1 per corner */
void l2_exchange(l2_world *w) {
    static const int face_pops[4][3] = {{1, 5, 8}, {3, 6, 7}, {2, 5, 6}, {4, 7, 8}};   /* to right, left, top, bottom */
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        for (int face = 0; face < 4; ++face) {
            if (R->nbr[face] < 0) continue;
            l2_rank *D = &w->r[R->nbr[face]];
            for (int s = 0; s < 3; ++s) {
                int a = face_pops[face][s];
                if (face < 2) for (int j = 1; j <= R->ny; ++j) FP(D, a, face == 0 ? 0 : D->nx + 1, j) = FP(R, a, face == 0 ? R->nx : 1, j);
                else for (int i = 1; i <= R->nx; ++i) FP(D, a, i, face == 2 ? 0 : D->ny + 1) = FP(R, a, i, face == 2 ? R->ny : 1);
            }
        }
        for (int a = 5; a < Q9; ++a) {
            if (R->cnr[a - 5] < 0) continue;
            l2_rank *D = &w->r[R->cnr[a - 5]];
            FP(D, a, ex[a] > 0 ? 0 : D->nx + 1, ey[a] > 0 ? 0 : D->ny + 1) = FP(R, a, ex[a] > 0 ? R->nx : 1, ey[a] > 0 ? R->ny : 1);
        }
    }
}

/* streaming(): evolution.f90:80-97 (pull).  Halo entries at physical walls are read as they are (the reference reads
 * them uninitialised) and overwritten by bounceback(); here they are whatever the caller left (0 after create). */
void l2_streaming(l2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i)
                for (int a = 0; a < Q9; ++a) F(R, a, i, j) = FP(R, a, i - ex[a], j - ey[a]);
    }
}

/* bounceback(): bounceback.f90:7-40 == boundary(), c:286-313 (left, right, bottom, then the moving top) */
void l2_bounceback(l2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        const int nx = R->nx, ny = R->ny;
        if (R->coords[0] == 0)
            for (int j = 1; j <= ny; ++j) { F(R, 1, 1, j) = FP(R, 3, 1, j); F(R, 5, 1, j) = FP(R, 7, 1, j); F(R, 8, 1, j) = FP(R, 6, 1, j); }
        if (R->coords[0] == w->dims[0] - 1)
            for (int j = 1; j <= ny; ++j) { F(R, 3, nx, j) = FP(R, 1, nx, j); F(R, 6, nx, j) = FP(R, 8, nx, j); F(R, 7, nx, j) = FP(R, 5, nx, j); }
        if (R->coords[1] == 0)
            for (int i = 1; i <= nx; ++i) { F(R, 2, i, 1) = FP(R, 4, i, 1); F(R, 5, i, 1) = FP(R, 7, i, 1); F(R, 6, i, 1) = FP(R, 8, i, 1); }
        if (R->coords[1] == w->dims[1] - 1)
            for (int i = 1; i <= nx; ++i) {
                F(R, 4, i, ny) = FP(R, 2, i, ny);
                if (w->variant == L2_I) {                /* inc:291-292 */
                    F(R, 7, i, ny) = FP(R, 5, i, ny) - (w->U0) / 6.0;
                    F(R, 8, i, ny) = FP(R, 6, i, ny) - (-w->U0) / 6.0;
                } else {
                    F(R, 7, i, ny) = FP(R, 5, i, ny) - S(R, rho, i, ny) * (w->U0) / 6.0;
                    F(R, 8, i, ny) = FP(R, 6, i, ny) - S(R, rho, i, ny) * (-w->U0) / 6.0;
                }
            }
    }
}

/* macro(): evolution.f90:107-113; c:321-336 accumulates the same terms in the same order (its f*0.0 terms add +-0) */
void l2_macro(l2_world *w) {
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        for (int j = 1; j <= R->ny; ++j)
            for (int i = 1; i <= R->nx; ++i) {
                const double *f = &F(R, 0, i, j);
                double rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
                S(R, rho, i, j) = rho;
                if (w->variant == L2_I) {                /* inc:307-308 */
                    S(R, u, i, j) = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]);
                    S(R, v, i, j) = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]);
                } else {
                    S(R, u, i, j) = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / rho;
                    S(R, v, i, j) = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / rho;
                }
            }
    }
}

/* check(): L2F evolution.f90:128-147 (rank sums, then Allreduce in rank order); L2C c:341-363 (pow(.,2), grouped add,
 * cells in i-outer / j-inner order, pow(.,0.5)); L2I inc:316-341 (sums of dsqrt, errorU = error1/error2) */
double l2_check(l2_world *w) {
    double t1 = 0.0, t2 = 0.0;
    for (int r = 0; r < w->np; ++r) {
        l2_rank *R = &w->r[r];
        double e1 = 0.0, e2 = 0.0;
        if (IS_C(w->variant)) {
            for (int i = 1; i <= R->nx; ++i)
                for (int j = 1; j <= R->ny; ++j) {
                    e1 += pow(S(R, u, i, j) - S(R, up, i, j), 2) + pow(S(R, v, i, j) - S(R, vp, i, j), 2);
                    e2 += pow(S(R, u, i, j), 2) + pow(S(R, v, i, j), 2);
                    S(R, up, i, j) = S(R, u, i, j); S(R, vp, i, j) = S(R, v, i, j);
                }
        } else if (w->variant == L2_I) {                 /* inc:325-332 */
            for (int j = 1; j <= R->ny; ++j)
                for (int i = 1; i <= R->nx; ++i) {
                    e1 = e1 + sqrt((S(R, u, i, j) - S(R, up, i, j)) * (S(R, u, i, j) - S(R, up, i, j)) + (S(R, v, i, j) - S(R, vp, i, j)) * (S(R, v, i, j) - S(R, vp, i, j)));
                    e2 = e2 + sqrt(S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j));
                    S(R, up, i, j) = S(R, u, i, j); S(R, vp, i, j) = S(R, v, i, j);
                }
        } else {
            for (int j = 1; j <= R->ny; ++j)
                for (int i = 1; i <= R->nx; ++i) {
                    e1 = e1 + (S(R, u, i, j) - S(R, up, i, j)) * (S(R, u, i, j) - S(R, up, i, j)) + (S(R, v, i, j) - S(R, vp, i, j)) * (S(R, v, i, j) - S(R, vp, i, j));
                    e2 = e2 + S(R, u, i, j) * S(R, u, i, j) + S(R, v, i, j) * S(R, v, i, j);
                }
            memcpy(R->up, R->u, sizeof(double) * (size_t)R->nx * R->ny);
            memcpy(R->vp, R->v, sizeof(double) * (size_t)R->nx * R->ny);
        }
        t1 += e1; t2 += e2;
    }
    w->errorU = IS_C(w->variant) ? pow(t1, 0.5) / pow(t2, 0.5) : w->variant == L2_I ? t1 / t2 : sqrt(t1) / sqrt(t2);   /* inc:335 */
    return w->errorU;
}

/* n iterations of the driver loop body: main.f90:66-82 == c:57-63 */
void l2_step(l2_world *w, int n) {
    for (int s = 0; s < n; ++s) {
        w->itc += 1;
        l2_collision(w);
        l2_exchange(w);
        l2_streaming(w);
        l2_bounceback(w);
        l2_macro(w);
    }
}
