// lbm_aa_kernels.inl -- D3Q19 lid-driven cavity on ONE lattice (AA-pattern): the reference loop body
// (L3/main.f90:89-97: collision, streaming, bounceback, macro) updated in place, so a GPU holds a lattice about 1.4x larger
// in cells than with the two ping-pong lattices of lbm_kernels.inl.  Compiled twice (lbm_aa.cu: namespace strict, -fmad=false;
// lbm_aa_fast.cu: namespace fast, -fmad=true); per-cell arithmetic is d3q19_mrt.inl's, identical to the ping-pong path.
//
// Two layouts of the same lattice A[slot][k][j][x] alternate:
//   NATURAL  A[a][x]      = f_a(x)            the populations as the reference holds them before collision()
//   POST     A[opp(a)][x] = f_post_a(x)       post-collision populations parked in the opposite slot of their own cell
// k_aa_even   NATURAL -> POST : macro() + collision() of one cell, all loads and stores at the cell itself.
// k_aa_odd    POST -> NATURAL : pull f_a(x) = A[opp(a)][x - e_a] (streaming() + bounceback() as the unified wall rule),
//                               macro(), collision(), then push f_post_a(x) to A[a][x + e_a] -- which is streaming() +
//                               bounceback() of the NEXT loop body -- so one launch advances two streaming steps.
// Every location is read and written by exactly one thread (the slot (s, z) is touched only by cell z - e_s, or by z itself at
// a wall), reads before writes: in place without races.  Either launch moves 19 loads + 19 stores = 304 B per cell.
// Aliasing note: the even and odd kernels read the lattice through a const __restrict__ alias with __ldg (non-coherent loads,
// free scheduling: 88-96 registers instead of 90-118).  That is sound here because a thread's stores all depend on all 19 of
// its loads (every post-collision population is a function of every moment), so no store can be hoisted above a load, and the
// only thread that ever writes a location during a launch is the one that read it: a stale non-coherent line can only be
// stale in entries nobody reads again.
#include "common.cuh"

namespace mglc {
namespace MGLC_NS {

#include "d3q19_mrt.inl"

struct AaWalls { bool xp, xm, yp, ym, zp, zm; };
__device__ __forceinline__ AaWalls aa_walls(const Geom &g, int i, int j, int k) {
    return AaWalls{g.wall[0] && i == g.nx, g.wall[1] && i == 1, g.wall[2] && j == g.ny,
                   g.wall[3] && j == 1,    g.wall[4] && k == g.nz, g.wall[5] && k == 1};
}
template <bool BGK>
__device__ __forceinline__ void aa_collide(const double (&f)[19], double rho, double u, double v, double w, const LbmParams &p,
                                           double (&fp)[19]) {
    if (BGK) d3q19_collide_bgk(f, rho, u, v, w, p.Snu, fp);
    else d3q19_collide(f, rho, u, v, w, p.Snu, p.Sq, fp);
}
// (a, opp(a), ex, ey, ez) of L3/commondata.f90:32-40
#define AA_FOR_ALL(X)                                                                                                   \
    X(1, 2, 1, 0, 0)    X(2, 1, -1, 0, 0)   X(3, 4, 0, 1, 0)    X(4, 3, 0, -1, 0)   X(5, 6, 0, 0, 1)    X(6, 5, 0, 0, -1)   \
    X(7, 10, 1, 1, 0)   X(8, 9, -1, 1, 0)   X(9, 8, 1, -1, 0)   X(10, 7, -1, -1, 0)                                         \
    X(11, 14, 1, 0, 1)  X(12, 13, -1, 0, 1) X(13, 12, 1, 0, -1) X(14, 11, -1, 0, -1)                                        \
    X(15, 18, 0, 1, 1)  X(16, 17, 0, -1, 1) X(17, 16, 0, 1, -1) X(18, 15, 0, -1, -1)

// POST layout -> the populations arriving at cell c: f_a(x) = f_post_a(x - e_a) = A[opp(a)][x - e_a]; if x - e_a lies
// outside the global box, bounceback() (L3/bounce_back.f90:6-83) gives f_a(x) = f_post_opp(a)(x) = A[a][x], minus the
// moving-lid term on populations 14 / 13 with rho of the previous macro() (:77-78).
#define AA_PULL(a, o, dx, dy, dz)                                                                           \
    {                                                                                                       \
        const bool wall_ = ((dx) == 1 && wf.xm) || ((dx) == -1 && wf.xp) || ((dy) == 1 && wf.ym) ||        \
                           ((dy) == -1 && wf.yp) || ((dz) == 1 && wf.zm) || ((dz) == -1 && wf.zp);         \
        f[a] = __ldg(Ain + (wall_ ? (a) * sq + c : (o) * sq + (c - (dz) * sz - (dy) * sy - (dx))));                    \
    }
#define AA_PULL_ALL()    \
    f[0] = __ldg(Ain + c); \
    AA_FOR_ALL(AA_PULL)
#define AA_LID_PULL(rho_lid)                                                                   \
    if (wf.zp) {                                                                               \
        const double r6 = __ddiv_rn((rho_lid)[(i - 1) + (long long)g.nx * (j - 1)], 6.0);     \
        f[14] = __dsub_rn(f[14], __dmul_rn(r6, p.U0));                                         \
        f[13] = __dsub_rn(f[13], __dmul_rn(r6, -p.U0));                                        \
    }
// NATURAL layout <- the post-collision populations of cell c: f_post_a travels to A[a][x + e_a]; if x + e_a lies outside the
// global box it comes back as population opp(a) of the cell itself, A[opp(a)][x] (bounceback() of the next loop body)
#define AA_PUSH(a, o, dx, dy, dz)                                                                           \
    {                                                                                                       \
        const bool out_ = ((dx) == 1 && wf.xp) || ((dx) == -1 && wf.xm) || ((dy) == 1 && wf.yp) ||         \
                          ((dy) == -1 && wf.ym) || ((dz) == 1 && wf.zp) || ((dz) == -1 && wf.zm);          \
        A[out_ ? (o) * sq + c : (a) * sq + (c + (dz) * sz + (dy) * sy + (dx))] = fp[a];                    \
    }

// prologue: collision() with the stored rho,u,v,w (what the reference does with the fields initial() or the caller left)
template <bool BGK>
__global__ void __launch_bounds__(128, 4) k_aa_collide0(Geom g, LbmParams p, double *A, const double *__restrict__ rho,
                                                        const double *__restrict__ u, const double *__restrict__ v,
                                                        const double *__restrict__ w) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = A[a * sq + c];
    aa_collide<BGK>(f, rho[m], u[m], v[m], w[m], p, fp);
    A[c] = fp[0];
#define AA_PARK(a, o, dx, dy, dz) A[(o) * sq + c] = fp[a];
    AA_FOR_ALL(AA_PARK)
}

// NATURAL -> POST: macro() of this loop body and collision() of the next, at the cell
template <bool BGK>
__global__ void __launch_bounds__(128, 5) k_aa_even(Geom g, LbmParams p, double *__restrict__ A, double *__restrict__ rho_lid_out) {
    const double *__restrict__ Ain = A;      // see the note on aliasing above
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k);
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = __ldg(Ain + a * sq + c);
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    aa_collide<BGK>(f, rho, u, v, w, p, fp);
    A[c] = fp[0];
    AA_FOR_ALL(AA_PARK)
#undef AA_PARK
    // the moving-lid bounce-back of the next streaming step needs this macro()'s rho on the lid plane (bounce_back.f90:77-78)
    if (g.lid && k == g.nz) rho_lid_out[(i - 1) + (long long)g.nx * (j - 1)] = rho;
}

// POST -> NATURAL: streaming() + bounceback() + macro() of loop body n, collision() + streaming() + bounceback() of body n+1
template <bool BGK>
__global__ void __launch_bounds__(128, 4) k_aa_odd(Geom g, LbmParams p, double *__restrict__ A, const double *__restrict__ rho_lid_in) {
    const double *__restrict__ Ain = A;      // see the note on aliasing above
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz, c = g.idx(0, i, j, k);
    const AaWalls wf = aa_walls(g, i, j, k);
    // Interior cells (no wall flag: all but the outermost shell) take straight-line paths: every address is the cell index plus
    // a block-uniform offset, no per-population wall test and no 64-bit select -- half the instructions of the general path.
    const bool shell = wf.xp | wf.xm | wf.yp | wf.ym | wf.zp | wf.zm;
    double f[19], fp[19];
    if (!shell) {
        f[0] = __ldg(Ain + c);
#define AA_PULL_IN(a, o, dx, dy, dz) f[a] = __ldg(Ain + ((o) * sq + (c - (dz) * sz - (dy) * sy - (dx))));
        AA_FOR_ALL(AA_PULL_IN)
#undef AA_PULL_IN
    } else {
        AA_PULL_ALL();
        AA_LID_PULL(rho_lid_in);
    }
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    aa_collide<BGK>(f, rho, u, v, w, p, fp);
    A[c] = fp[0];
    if (!shell) {
#define AA_PUSH_IN(a, o, dx, dy, dz) A[(a) * sq + (c + (dz) * sz + (dy) * sy + (dx))] = fp[a];
        AA_FOR_ALL(AA_PUSH_IN)
#undef AA_PUSH_IN
    } else {
        if (wf.zp) {  // the lid term of body n+1 uses the rho this macro() just produced: f(14) = f_post(11) - rho/6*U0, f(13) = f_post(12) - rho/6*(-U0)
            const double r6 = __ddiv_rn(rho, 6.0);
            fp[11] = __dsub_rn(fp[11], __dmul_rn(r6, p.U0));
            fp[12] = __dsub_rn(fp[12], __dmul_rn(r6, -p.U0));
        }
        AA_FOR_ALL(AA_PUSH)
    }
}

#ifndef MGLC_HOST_SHIM
static inline dim3 aa_grid(const Geom &g) { return dim3((g.nx + 127) / 128, g.ny, g.nz); }
int launch_aa_collide0(const Geom &g, const LbmParams &p, double *A, const double *rho, const double *u, const double *v, const double *w,
                       cudaStream_t s) {
    if (p.bgk) k_aa_collide0<true><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho, u, v, w);
    else k_aa_collide0<false><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho, u, v, w);
    return 1;
}
int launch_aa_even(const Geom &g, const LbmParams &p, double *A, double *rho_lid_out, cudaStream_t s) {
    if (p.bgk) k_aa_even<true><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho_lid_out);
    else k_aa_even<false><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho_lid_out);
    return 1;
}
int launch_aa_odd(const Geom &g, const LbmParams &p, double *A, const double *rho_lid_in, cudaStream_t s) {
    if (p.bgk) k_aa_odd<true><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho_lid_in);
    else k_aa_odd<false><<<aa_grid(g), 128, 0, s>>>(g, p, A, rho_lid_in);
    return 1;
}
#endif

}  // namespace MGLC_NS
}  // namespace mglc
