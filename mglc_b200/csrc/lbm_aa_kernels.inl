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
// Decomposed lattices (PEER = true, the neighbours' lattices mapped: peer pointers or CUDA IPC; the PeerTable of common.cuh is a
// kernel parameter here, read from the constant bank, so every global load of these kernels is a lattice load).  The two
// exchanges of a decomposed AA run ride along with the launches, like the halo stores of k_fused<..., PEER>:
//   * a launch that PARKS (k_aa_collide0, k_aa_even) also stores the parked populations of its boundary cells into the
//     neighbour's halo cells, same slot opp(a) -- the message sets of ex_sendrecv.f90:12-123 (5 per face, 1 per edge) -- so the
//     next odd launch over there can pull them;
//   * the odd launch PUSHES f_post_a(x) to x + e_a; where that cell belongs to a neighbour the store goes straight into the
//     neighbour's own cell (slot a), so no halo layer has to be sent back.
// Nobody reads a location that another subdomain writes during the same launch: (a, y) of a boundary cell y is pulled by the
// cell x = y - e_a across the face -- from x's own halo copy -- and pushed by that same cell; y's owner only touches it in the
// next launch.  The launches of neighbouring subdomains are ordered by the neighbour barrier (events / flag words).
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
// Cache hints and occupancy were measured on the two bulk kernels (640^3, 100 steps, profiles/r2q_aa_memop_640.txt): loads as
// ld.global.lu / .cs 1.9 % slower than .nc, stores as st.global.cs / .wt 9-11 % slower than plain stores (the L2 merges the 19
// store streams into full lines before writing back), 6 CTAs/SM (80 registers) 5.9 % slower, 4 CTAs/SM (no spill in
// k_aa_even) 1 % slower.  So: non-coherent loads, plain stores, 5 CTAs/SM.
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

// ---- decomposed lattices: where the parked populations of a boundary cell also go, and where a push across a face lands ----
// halo cell of the neighbour in direction (EX,EY,EZ) that mirrors cell (i,j,k) (as peer_cell in lbm_kernels.inl)
template <int EX, int EY, int EZ>
__device__ __forceinline__ long long aa_peer_halo(const PeerTable &pt, int D, int i, int j, int k) {
    const int ii = EX > 0 ? 0 : (EX < 0 ? pt.n[D][0] + 1 : i);
    const int jj = EY > 0 ? 0 : (EY < 0 ? pt.n[D][1] + 1 : j);
    const int kk = EZ > 0 ? 0 : (EZ < 0 ? pt.n[D][2] + 1 : k);
    return kk * pt.sz[D] + jj * pt.sy[D] + (ii + OX - 1);
}
// population a of a face message is parked in slot opp(a)
#define AA_PARK_FACE(D, EX, EY, EZ, a0, o0, a1, o1, a2, o2, a3, o3, a4, o4)                                   \
    if (pm & (1u << (D))) {                                                                                   \
        double *P = pt.F[D];                                                                                 \
        const long long hc = aa_peer_halo<EX, EY, EZ>(pt, D, i, j, k), q = pt.sq[D];                         \
        P[(o0) * q + hc] = fp[a0]; P[(o1) * q + hc] = fp[a1]; P[(o2) * q + hc] = fp[a2];                      \
        P[(o3) * q + hc] = fp[a3]; P[(o4) * q + hc] = fp[a4];                                                 \
    }
#define AA_PARK_EDGE(A, O, EX, EY, EZ)                                                                       \
    if (pm & (1u << (A))) pt.F[A][(O) * pt.sq[A] + aa_peer_halo<EX, EY, EZ>(pt, A, i, j, k)] = fp[A];
// a cell on a face of the block may store into a neighbour -- unless the neighbour barrier has failed (sticky error word of
// k_halo_wait, one process per GPU only): from then on nothing goes into a neighbour any more, as on the two-lattice path.
// Read before the cell's lattice loads, so that no global load follows a store.
__device__ __forceinline__ bool aa_may_store_to_peers(const PeerTable &pt, const Geom &g, int i, int j, int k) {
    if (!((i == 1) | (i == g.nx) | (j == 1) | (j == g.ny) | (k == 1) | (k == g.nz))) return false;
    return !(pt.err && *(const volatile int *)pt.err);
}
__device__ __forceinline__ void aa_peer_park(const PeerTable &pt, const Geom &g, int i, int j, int k, const double (&fp)[19]) {
    const unsigned pm = pt.mask;
    const bool xp = i == g.nx, xm = i == 1, yp = j == g.ny, ym = j == 1, zp = k == g.nz, zm = k == 1;
    if (!(xp | xm | yp | ym | zp | zm)) return;
    if (xp) AA_PARK_FACE(0, 1, 0, 0, 1, 2, 7, 10, 9, 8, 11, 14, 13, 12)
    if (xm) AA_PARK_FACE(1, -1, 0, 0, 2, 1, 8, 9, 10, 7, 12, 13, 14, 11)
    if (yp) AA_PARK_FACE(2, 0, 1, 0, 3, 4, 7, 10, 8, 9, 15, 18, 17, 16)
    if (ym) AA_PARK_FACE(3, 0, -1, 0, 4, 3, 9, 8, 10, 7, 16, 17, 18, 15)
    if (zp) AA_PARK_FACE(4, 0, 0, 1, 5, 6, 11, 14, 12, 13, 15, 18, 16, 17)
    if (zm) AA_PARK_FACE(5, 0, 0, -1, 6, 5, 13, 12, 14, 11, 17, 16, 18, 15)
    if (xp && yp) AA_PARK_EDGE(7, 10, 1, 1, 0)
    if (xm && yp) AA_PARK_EDGE(8, 9, -1, 1, 0)
    if (xp && ym) AA_PARK_EDGE(9, 8, 1, -1, 0)
    if (xm && ym) AA_PARK_EDGE(10, 7, -1, -1, 0)
    if (xp && zp) AA_PARK_EDGE(11, 14, 1, 0, 1)
    if (xm && zp) AA_PARK_EDGE(12, 13, -1, 0, 1)
    if (xp && zm) AA_PARK_EDGE(13, 12, 1, 0, -1)
    if (xm && zm) AA_PARK_EDGE(14, 11, -1, 0, -1)
    if (yp && zp) AA_PARK_EDGE(15, 18, 0, 1, 1)
    if (ym && zp) AA_PARK_EDGE(16, 17, 0, -1, 1)
    if (yp && zm) AA_PARK_EDGE(17, 16, 0, 1, -1)
    if (ym && zm) AA_PARK_EDGE(18, 15, 0, -1, -1)
}
// push of population a = (dx,dy,dz) from cell (i,j,k) of a decomposed lattice: beyond a wall of the global box it comes back as
// opp(a) of the cell (bounceback(), as AA_PUSH); beyond a face shared with a neighbour it is stored into that neighbour's own
// cell -- the face neighbour if one component leaves the block, the edge neighbour (direction = a itself) if two do
#define AA_PUSH_PEER(a, o, dx, dy, dz)                                                                                  \
    {                                                                                                                   \
        const bool out_ = ((dx) == 1 && wf.xp) || ((dx) == -1 && wf.xm) || ((dy) == 1 && wf.yp) ||                     \
                          ((dy) == -1 && wf.ym) || ((dz) == 1 && wf.zp) || ((dz) == -1 && wf.zm);                      \
        const bool bx = ((dx) == 1 && i == g.nx) || ((dx) == -1 && i == 1);                                            \
        const bool by = ((dy) == 1 && j == g.ny) || ((dy) == -1 && j == 1);                                            \
        const bool bz = ((dz) == 1 && k == g.nz) || ((dz) == -1 && k == 1);                                            \
        if (out_) A[(o) * sq + c] = fp[a];                                                                             \
        else if (!(bx | by | bz)) A[(a) * sq + (c + (dz) * sz + (dy) * sy + (dx))] = fp[a];                            \
        else {                                                                                                          \
            const int D = (int)bx + (int)by + (int)bz == 2 ? (a) : bx ? ((dx) > 0 ? 0 : 1) : by ? ((dy) > 0 ? 2 : 3) : ((dz) > 0 ? 4 : 5); \
            const int ii = bx ? ((dx) > 0 ? 1 : pt.n[D][0]) : i + (dx);                                               \
            const int jj = by ? ((dy) > 0 ? 1 : pt.n[D][1]) : j + (dy);                                               \
            const int kk = bz ? ((dz) > 0 ? 1 : pt.n[D][2]) : k + (dz);                                               \
            pt.F[D][(a) * pt.sq[D] + kk * pt.sz[D] + jj * pt.sy[D] + (ii + OX - 1)] = fp[a];                       \
        }                                                                                                               \
    }

// prologue: collision() with the stored rho,u,v,w (what the reference does with the fields initial() or the caller left)
template <bool BGK, bool PEER>
__global__ void __launch_bounds__(128, 4) k_aa_collide0(Geom g, LbmParams p, double *A, const double *__restrict__ rho,
                                                        const double *__restrict__ u, const double *__restrict__ v,
                                                        const double *__restrict__ w, const __grid_constant__ PeerTable pt) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    const bool park = PEER && aa_may_store_to_peers(pt, g, i, j, k);
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = A[a * sq + c];
    aa_collide<BGK>(f, rho[m], u[m], v[m], w[m], p, fp);
    A[c] = fp[0];
#define AA_PARK(a, o, dx, dy, dz) A[(o) * sq + c] = fp[a];
    AA_FOR_ALL(AA_PARK)
    if (PEER && park) aa_peer_park(pt, g, i, j, k, fp);
}

// NATURAL -> POST: macro() of this loop body and collision() of the next, at the cell
template <bool BGK, bool PEER>
__global__ void __launch_bounds__(128, 5) k_aa_even(Geom g, LbmParams p, double *__restrict__ A, double *__restrict__ rho_lid_out,
                                                    const __grid_constant__ PeerTable pt) {
    const double *__restrict__ Ain = A;      // see the note on aliasing above
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k);
    const bool park = PEER && aa_may_store_to_peers(pt, g, i, j, k);
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = __ldg(Ain + a * sq + c);
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    aa_collide<BGK>(f, rho, u, v, w, p, fp);
    A[c] = fp[0];
    AA_FOR_ALL(AA_PARK)
#undef AA_PARK
    if (PEER && park) aa_peer_park(pt, g, i, j, k, fp);
    // the moving-lid bounce-back of the next streaming step needs this macro()'s rho on the lid plane (bounce_back.f90:77-78)
    if (g.lid && k == g.nz) rho_lid_out[(i - 1) + (long long)g.nx * (j - 1)] = rho;
}

// POST -> NATURAL: streaming() + bounceback() + macro() of loop body n, collision() + streaming() + bounceback() of body n+1
template <bool BGK, bool PEER>
__global__ void __launch_bounds__(128, 4) k_aa_odd(Geom g, LbmParams p, double *__restrict__ A, const double *__restrict__ rho_lid_in,
                                                   const __grid_constant__ PeerTable pt) {
    const double *__restrict__ Ain = A;      // see the note on aliasing above
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz, c = g.idx(0, i, j, k);
    const AaWalls wf = aa_walls(g, i, j, k);
    // Three paths, chosen per WARP (j, k are block-uniform, and whether a warp holds the first or the last cell of a row follows
    // from its first index), so that no warp runs two of them:
    //   0  interior: every address is the cell index plus a block-uniform offset, no wall test, no 64-bit select;
    //   1  the row is interior in y and z, the warp holds an x wall: the ten populations with an x component choose between
    //      the neighbour and the bounce-back slot of the cell itself (one select per address); the other nine as path 0.
    //      (Per-lane branching here made one warp in twelve run path 0 AND the general path: 0.5 ms of a 22.4 ms step.)
    //   2  general: rows on a y or z face of the block, and -- decomposed blocks -- warps holding a cell of an x face shared
    //      with a neighbour; a cell on a shared face pushes through AA_PUSH_PEER, any other cell as in a one-block run.
    const int i0 = 1 + (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
    const bool warp_lo = i0 == 1, warp_hi = (g.nx >= i0) & (g.nx - i0 < 32);
    const bool yz_face = (j == 1) | (j == g.ny) | (k == 1) | (k == g.nz);
    const bool x_shared = PEER && ((warp_lo && !g.wall[1]) | (warp_hi && !g.wall[0]));
    const int path = (yz_face | x_shared) ? 2 : ((warp_lo | warp_hi) ? 1 : 0);
    const bool peer_cell = PEER && ((i == g.nx && !g.wall[0]) | (i == 1 && !g.wall[1]) | (j == g.ny && !g.wall[2]) |
                                    (j == 1 && !g.wall[3]) | (k == g.nz && !g.wall[4]) | (k == 1 && !g.wall[5])) &&
                           aa_may_store_to_peers(pt, g, i, j, k);     // barrier failed: the push stays in the block's own halo ring
    double f[19], fp[19];
    if (path == 0) {
        f[0] = __ldg(Ain + c);
#define AA_PULL_IN(a, o, dx, dy, dz) f[a] = __ldg(Ain + ((o) * sq + (c - (dz) * sz - (dy) * sy - (dx))));
        AA_FOR_ALL(AA_PULL_IN)
#undef AA_PULL_IN
    } else if (path == 1) {
        f[0] = __ldg(Ain + c);
#define AA_PULL_X(a, o, dx, dy, dz)                                                                                  \
    {                                                                                                                \
        const bool wall_ = ((dx) == 1 && wf.xm) || ((dx) == -1 && wf.xp);                                           \
        f[a] = __ldg(Ain + (wall_ ? (a) * sq + c : (o) * sq + (c - (dz) * sz - (dy) * sy - (dx))));                  \
    }
        AA_FOR_ALL(AA_PULL_X)
#undef AA_PULL_X
    } else {
        AA_PULL_ALL();
        AA_LID_PULL(rho_lid_in);
    }
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    aa_collide<BGK>(f, rho, u, v, w, p, fp);
    A[c] = fp[0];
    if (path == 0) {
#define AA_PUSH_IN(a, o, dx, dy, dz) A[(a) * sq + (c + (dz) * sz + (dy) * sy + (dx))] = fp[a];
        AA_FOR_ALL(AA_PUSH_IN)
#undef AA_PUSH_IN
    } else if (path == 1) {
#define AA_PUSH_X(a, o, dx, dy, dz)                                                                                  \
    {                                                                                                                \
        const bool out_ = ((dx) == 1 && wf.xp) || ((dx) == -1 && wf.xm);                                            \
        A[out_ ? (o) * sq + c : (a) * sq + (c + (dz) * sz + (dy) * sy + (dx))] = fp[a];                             \
    }
        AA_FOR_ALL(AA_PUSH_X)
#undef AA_PUSH_X
    } else {
        if (wf.zp) {  // the lid term of body n+1 uses the rho this macro() just produced: f(14) = f_post(11) - rho/6*U0, f(13) = f_post(12) - rho/6*(-U0)
            const double r6 = __ddiv_rn(rho, 6.0);
            fp[11] = __dsub_rn(fp[11], __dmul_rn(r6, p.U0));
            fp[12] = __dsub_rn(fp[12], __dmul_rn(r6, -p.U0));
        }
        if (PEER && peer_cell) { AA_FOR_ALL(AA_PUSH_PEER) }
        else { AA_FOR_ALL(AA_PUSH) }
    }
}

#ifndef MGLC_HOST_SHIM
static const PeerTable aa_no_peers = {};
static inline dim3 aa_grid(const Geom &g) { return dim3((g.nx + 127) / 128, g.ny, g.nz); }
#define MGLC_AA_LAUNCH(K, ...)                                                                          \
    do {                                                                                                \
        if (pt) { if (p.bgk) K<true, true><<<aa_grid(g), 128, 0, s>>>(__VA_ARGS__, *pt); else K<false, true><<<aa_grid(g), 128, 0, s>>>(__VA_ARGS__, *pt); } \
        else { if (p.bgk) K<true, false><<<aa_grid(g), 128, 0, s>>>(__VA_ARGS__, aa_no_peers); else K<false, false><<<aa_grid(g), 128, 0, s>>>(__VA_ARGS__, aa_no_peers); } \
    } while (0)
int launch_aa_collide0(const Geom &g, const LbmParams &p, double *A, const double *rho, const double *u, const double *v, const double *w,
                       cudaStream_t s, const PeerTable *pt) {
    MGLC_AA_LAUNCH(k_aa_collide0, g, p, A, rho, u, v, w);
    return 1;
}
int launch_aa_even(const Geom &g, const LbmParams &p, double *A, double *rho_lid_out, cudaStream_t s, const PeerTable *pt) {
    MGLC_AA_LAUNCH(k_aa_even, g, p, A, rho_lid_out);
    return 1;
}
int launch_aa_odd(const Geom &g, const LbmParams &p, double *A, const double *rho_lid_in, cudaStream_t s, const PeerTable *pt) {
    MGLC_AA_LAUNCH(k_aa_odd, g, p, A, rho_lid_in);
    return 1;
}
#undef MGLC_AA_LAUNCH
#endif

}  // namespace MGLC_NS
}  // namespace mglc
