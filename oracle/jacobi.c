/*
 * oracle/jacobi.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's Jacobi
 * halo-exchange path (cheryli/MGLC, MPI/Laplace/fortran/jacobi2d_mpi.f90, "LAP" below) and its 3-D
 * extension.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 *
 * PARITY PIN: the 2-D mode is checked bit for bit against the reference's own compiled C program
 * MPI/Laplace/c/laplace2d.c (functions jacobi() and swap(), built from the source where it lies into
 * oracle/_ref/liblaplace2d_ref.so by `make -C oracle ref`), and against the Fortran program's own text
 * (jacobi LAP:176-180 with a source term, check_diff LAP:193-198, init LAP:151-165, machine-evaluated by
 * tests/golden/make_golden_jacobi_fortran.py; and the program's loop LAP:91-112 as a whole: 300 iterations of two sweeps with
 * check_diff every 100, reproduced on 1, 4 and 6 emulated ranks), see tests/test_oracle_jacobi.py.  The 3-D
 * mode has no reference (LAP is 2-D only; BASELINE.json config 2 asks for 512^3): it is defined here
 * as the same update with the six face neighbours, summed x-,x+,y-,y+,z-,z+ then + f, times the
 * compile-time constant 1/6 -- the 2-D mode is its nz = 1 special case with 0.25.
 *
 * Layout is the reference's: Fortran column-major A(0:nx+1, 0:ny+1 [, 0:nz+1]) with one ghost layer
 * (LAP:78-81).  One process emulates all P MPI ranks of the Cartesian grid (LAP:45-64).
 * Build with -ffp-contract=off: every operation is one IEEE fp64 rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct jac_rank {
    int n[3];            /* local nx, ny, nz (nz = 1 in 2-D) */
    int coords[3], start[3];
    int nbr[6];          /* +x -x +y -y +z -z, -1 = MPI_PROC_NULL (LAP:56-57) */
    double *A, *A_new, *f, *A_p;
} jac_rank;

typedef struct jac_world {
    int ndim, total[3], dims[3], np;
    jac_rank *r;
} jac_world;

static size_t sx(const jac_rank *R) { return (size_t)R->n[0] + 2; }
static size_t sy(const jac_rank *R) { return (size_t)R->n[1] + 2; }
static size_t count(const jac_world *W, const jac_rank *R) {
    return sx(R) * sy(R) * (W->ndim == 3 ? (size_t)R->n[2] + 2 : 1);
}
/* 2-D arrays have no k index: IDX(i,j,0) */
#define IDX(R, i, j, k) ((size_t)(i) + sx(R) * ((size_t)(j) + sy(R) * (size_t)(k)))

/* MPI_Dims_create(np, ndim, dims) with dims = 0 (LAP:47): balanced, non-increasing */
void jac_dims_create(int np, int ndim, int dims[3]) {
    int best[3] = {np, 1, 1};
    for (int a = 1; a <= np; ++a) {
        if (np % a) continue;
        for (int b = 1; b <= a; ++b) {
            if ((np / a) % b) continue;
            int c = np / a / b;
            if (c > b) continue;
            if (ndim == 2 && c != 1) continue;
            if (a < best[0] || (a == best[0] && b < best[1])) { best[0] = a; best[1] = b; best[2] = c; }
        }
    }
    memcpy(dims, best, sizeof best);
}

/* decompose_1d, LAP:130-141 (+ the global offset) */
static void decompose_1d(int total_n, int rank, int np, int *local_n, int *start) {
    int n = total_n / np, m = total_n % np;
    *local_n = n + (rank < m ? 1 : 0);
    *start = rank * n + (rank < m ? rank : m);
}

static int cart_rank(const jac_world *W, int c0, int c1, int c2) {
    if (c0 < 0 || c0 >= W->dims[0] || c1 < 0 || c1 >= W->dims[1] || c2 < 0 || c2 >= W->dims[2]) return -1;
    return (c0 * W->dims[1] + c1) * W->dims[2] + c2;
}

jac_world *jac_world_create(int ndim, int nx, int ny, int nz, int np, const int *dims_or_zero) {
    jac_world *W = calloc(1, sizeof *W);
    W->ndim = ndim; W->np = np;
    W->total[0] = nx; W->total[1] = ny; W->total[2] = ndim == 3 ? nz : 1;
    if (dims_or_zero && dims_or_zero[0] > 0) memcpy(W->dims, dims_or_zero, sizeof W->dims);
    else jac_dims_create(np, ndim, W->dims);
    W->r = calloc((size_t)np, sizeof(jac_rank));
    for (int r = 0; r < np; ++r) {
        jac_rank *R = &W->r[r];
        R->coords[2] = r % W->dims[2];
        R->coords[1] = (r / W->dims[2]) % W->dims[1];
        R->coords[0] = r / (W->dims[2] * W->dims[1]);
        for (int d = 0; d < 3; ++d) decompose_1d(W->total[d], R->coords[d], W->dims[d], &R->n[d], &R->start[d]);
        for (int d = 0; d < 3; ++d) {
            int p[3] = {R->coords[0], R->coords[1], R->coords[2]}, m[3] = {R->coords[0], R->coords[1], R->coords[2]};
            p[d] += 1; m[d] -= 1;
            R->nbr[2 * d] = cart_rank(W, p[0], p[1], p[2]);
            R->nbr[2 * d + 1] = cart_rank(W, m[0], m[1], m[2]);
        }
        size_t n = count(W, R);
        R->A = calloc(n, sizeof(double)); R->A_new = calloc(n, sizeof(double));
        R->f = calloc(n, sizeof(double)); R->A_p = calloc(n, sizeof(double));
    }
    return W;
}

void jac_world_destroy(jac_world *W) {
    if (!W) return;
    for (int r = 0; r < W->np; ++r) { free(W->r[r].A); free(W->r[r].A_new); free(W->r[r].f); free(W->r[r].A_p); }
    free(W->r); free(W);
}

/* which: 0 A, 1 A_new, 2 f, 3 A_p */
double *jac_ptr(jac_world *W, int r, int which) {
    jac_rank *R = &W->r[r];
    return which == 0 ? R->A : which == 1 ? R->A_new : which == 2 ? R->f : R->A_p;
}
void jac_rank_info(jac_world *W, int r, int *out /*[15]*/) {
    jac_rank *R = &W->r[r];
    memcpy(out, R->n, 12); memcpy(out + 3, R->coords, 12); memcpy(out + 6, R->start, 12); memcpy(out + 9, R->nbr, 24);
}
void jac_world_dims(jac_world *W, int *dims) { memcpy(dims, W->dims, 12); }

/* init(), LAP:144-166: A = A_new = f = 0; the top ghost layer (j = ny+1 in 2-D, k = nz+1 in 3-D) of the
 * ranks on the top of the process grid = 1, in both A and A_new, rims included (i = 0..nx+1).  A_p = A (LAP:92). */
void jac_init(jac_world *W) {
    for (int r = 0; r < W->np; ++r) {
        jac_rank *R = &W->r[r];
        size_t n = count(W, R);
        memset(R->A, 0, n * 8); memset(R->A_new, 0, n * 8); memset(R->f, 0, n * 8);
        int top = W->ndim - 1;
        if (R->coords[top] == W->dims[top] - 1) {
            if (W->ndim == 2) {
                for (int i = 0; i <= R->n[0] + 1; ++i) { R->A[IDX(R, i, R->n[1] + 1, 0)] = 1.0; R->A_new[IDX(R, i, R->n[1] + 1, 0)] = 1.0; }
            } else {
                for (int j = 0; j <= R->n[1] + 1; ++j)
                    for (int i = 0; i <= R->n[0] + 1; ++i) { R->A[IDX(R, i, j, R->n[2] + 1)] = 1.0; R->A_new[IDX(R, i, j, R->n[2] + 1)] = 1.0; }
            }
        }
        memcpy(R->A_p, R->A, n * 8);
    }
}

/* exchange_message(A), LAP:223-254: the last interior layer of each face goes to the neighbour's
 * opposite ghost layer; interior ranges only ("corners don't matter", LAP:231).  NOTE: the
 * reference's column type spans ny+2 entries starting at j = 1 (LAP:84, 245-252), i.e. it also moves
 * the j = ny+1 rim entry and one entry past the array; neither is ever read by the 5-point update,
 * so only j = 1..ny is moved here. */
void jac_exchange(jac_world *W) {
    for (int r = 0; r < W->np; ++r) {
        jac_rank *S = &W->r[r];
        const int k0 = W->ndim == 3 ? 1 : 0, k1 = W->ndim == 3 ? S->n[2] : 0;
        for (int face = 0; face < 2 * W->ndim; ++face) {
            if (S->nbr[face] < 0) continue;
            jac_rank *D = &W->r[S->nbr[face]];
            const int axis = face >> 1, plus = !(face & 1);
            if (axis == 0) {
                const int is = plus ? S->n[0] : 1, id = plus ? 0 : D->n[0] + 1;
                for (int k = k0; k <= k1; ++k) for (int j = 1; j <= S->n[1]; ++j) D->A[IDX(D, id, j, k)] = S->A[IDX(S, is, j, k)];
            } else if (axis == 1) {
                const int js = plus ? S->n[1] : 1, jd = plus ? 0 : D->n[1] + 1;
                for (int k = k0; k <= k1; ++k) for (int i = 1; i <= S->n[0]; ++i) D->A[IDX(D, i, jd, k)] = S->A[IDX(S, i, js, k)];
            } else {
                const int ks = plus ? S->n[2] : 1, kd = plus ? 0 : D->n[2] + 1;
                for (int j = 1; j <= S->n[1]; ++j) for (int i = 1; i <= S->n[0]; ++i) D->A[IDX(D, i, j, kd)] = S->A[IDX(S, i, j, ks)];
            }
        }
    }
}

/* jacobi(A, A_new), LAP:170-182, then the caller's role swap (LAP:97-103 ping-pongs A and A_new) */
void jac_sweep(jac_world *W) {
    for (int r = 0; r < W->np; ++r) {
        jac_rank *R = &W->r[r];
        const double *A = R->A, *f = R->f;
        double *B = R->A_new;
        if (W->ndim == 2) {
#pragma omp parallel for schedule(static)
            for (int j = 1; j <= R->n[1]; ++j)
                for (int i = 1; i <= R->n[0]; ++i)
                    B[IDX(R, i, j, 0)] = 0.25 * (A[IDX(R, i - 1, j, 0)] + A[IDX(R, i + 1, j, 0)] + A[IDX(R, i, j - 1, 0)] +
                                                 A[IDX(R, i, j + 1, 0)] + f[IDX(R, i, j, 0)]);
        } else {
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= R->n[2]; ++k)
                for (int j = 1; j <= R->n[1]; ++j)
                    for (int i = 1; i <= R->n[0]; ++i)
                        B[IDX(R, i, j, k)] = (1.0 / 6.0) * (A[IDX(R, i - 1, j, k)] + A[IDX(R, i + 1, j, k)] + A[IDX(R, i, j - 1, k)] +
                                                            A[IDX(R, i, j + 1, k)] + A[IDX(R, i, j, k - 1)] + A[IDX(R, i, j, k + 1)] +
                                                            f[IDX(R, i, j, k)]);
        }
        double *t = R->A; R->A = R->A_new; R->A_new = t;
    }
}

/* nits x { exchange_message(A); jacobi(A -> A_new); swap roles }  (LAP:94-103 does two per pass) */
void jac_step(jac_world *W, int nits) {
    for (int it = 0; it < nits; ++it) { jac_exchange(W); jac_sweep(W); }
}

/* check_diff + MPI_Allreduce(MAX), LAP:105-107,185-204: max |A_p - A| over the interior, then A_p = A */
double jac_check_diff(jac_world *W) {
    double error_max = 0.0;
    for (int r = 0; r < W->np; ++r) {
        jac_rank *R = &W->r[r];
        const int k0 = W->ndim == 3 ? 1 : 0, k1 = W->ndim == 3 ? R->n[2] : 0;
        double error = 0.0;
        for (int k = k0; k <= k1; ++k)
            for (int j = 1; j <= R->n[1]; ++j)
                for (int i = 1; i <= R->n[0]; ++i) {
                    double d = fabs(R->A_p[IDX(R, i, j, k)] - R->A[IDX(R, i, j, k)]);
                    error = error > d ? error : d;
                }
        memcpy(R->A_p, R->A, count(W, R) * 8);
        error_max = error_max > error ? error_max : error;
    }
    return error_max;
}
