// lbm_kernels.inl -- D3Q19 lattice-update kernels; compiled twice (lbm_strict.cu / lbm_fast.cu).
//
// Device layout (common.cuh): SoA F[a][k][j][x] with a one-cell halo; thread <-> cell, threadIdx.x runs
// along x so every population access of a warp is one contiguous 256-byte segment.  The pull scheme
// reads population a at (x - e_a): reads with ex = +-1 are shifted by 8 bytes (served from the same
// L1/L2 lines), every store is 128-byte aligned.
//
// The fused kernel is the reference loop body rotated by half a step: it performs streaming()
// (L3/streaming.f90) + bounceback() (L3/bounce_back.f90, folded into the pull, see d3q19_walls) +
// macro() (L3/macro.f90) of step n and collision() (L3/collision.f90) of step n+1, so a step costs
// one read and one write of the 19 populations: 304 B/cell.
#include "common.cuh"

namespace mglc {
namespace MGLC_NS {

#include "d3q19_mrt.inl"
#include "d3q19_thermal.inl"

// Pull the 19 populations that arrive at cell c (c = linear index of the cell in population 0), with
// bounceback() (L3/bounce_back.f90:6-83) folded in as the unified boundary rule:
// a population whose upstream cell x - e_a lies outside the GLOBAL box takes the opposite
// post-collision population of the cell itself, f_a(x) = f_post_opp(a)(x); through the moving lid,
// populations 14 / 13 also get - rho/6*(+U0) / - rho/6*(-U0) with rho left by the previous macro()
// (:77-78).  The value is the same whichever wall "wins", and z-last precedence is kept because the lid
// term is keyed on the cell.  The wall test only selects the load ADDRESS (no divergent branch, no
// extra live registers); wall halos are never read.
struct WallFlags { bool xp, xm, yp, ym, zp, zm; };
__device__ __forceinline__ WallFlags wall_flags(const Geom &g, int i, int j, int k) {
    return WallFlags{g.wall[0] && i == g.nx, g.wall[1] && i == 1, g.wall[2] && j == g.ny,
                     g.wall[3] && j == 1,    g.wall[4] && k == g.nz, g.wall[5] && k == 1};
}
#define MGLC_PULL(a, o, dx, dy, dz)                                                                         \
    {                                                                                                       \
        const bool wall_ = ((dx) == 1 && wf.xm) || ((dx) == -1 && wf.xp) || ((dy) == 1 && wf.ym) ||        \
                           ((dy) == -1 && wf.yp) || ((dz) == 1 && wf.zm) || ((dz) == -1 && wf.zp);         \
        f[a] = __ldg(Fin + (wall_ ? (o) * sq + c : (a) * sq + (c - (dz) * sz - (dy) * sy - (dx))));        \
    }
#define MGLC_PULL_ALL()                                                                                     \
    f[0] = __ldg(Fin + c);                                                                                  \
    MGLC_PULL(1, 2, 1, 0, 0)    MGLC_PULL(2, 1, -1, 0, 0)   MGLC_PULL(3, 4, 0, 1, 0)    MGLC_PULL(4, 3, 0, -1, 0)   \
    MGLC_PULL(5, 6, 0, 0, 1)    MGLC_PULL(6, 5, 0, 0, -1)                                                   \
    MGLC_PULL(7, 10, 1, 1, 0)   MGLC_PULL(8, 9, -1, 1, 0)   MGLC_PULL(9, 8, 1, -1, 0)   MGLC_PULL(10, 7, -1, -1, 0) \
    MGLC_PULL(11, 14, 1, 0, 1)  MGLC_PULL(12, 13, -1, 0, 1) MGLC_PULL(13, 12, 1, 0, -1) MGLC_PULL(14, 11, -1, 0, -1) \
    MGLC_PULL(15, 18, 0, 1, 1)  MGLC_PULL(16, 17, 0, -1, 1) MGLC_PULL(17, 16, 0, 1, -1) MGLC_PULL(18, 15, 0, -1, -1)
// moving lid; explicit _rn intrinsics so the fast build cannot contract this into an FMA
// rho_lid: the lid plane (k = nz) of rho as the previous macro() left it, indexed (i-1) + nx*(j-1)
#define MGLC_LID(rho_lid)                                                         \
    if (wf.zp) {                                                                  \
        const double r6 = __ddiv_rn((rho_lid)[(i - 1) + (long long)g.nx * (j - 1)], 6.0); \
        f[14] = __dsub_rn(f[14], __dmul_rn(r6, p.U0));                            \
        f[13] = __dsub_rn(f[13], __dmul_rn(r6, -p.U0));                           \
    }


// ---- direct halo stores (PeerTable, common.cuh) ------------------------------------------------------------------
// Halo cell of the neighbour in direction (EX,EY,EZ) that receives what cell (i,j,k) sends there: along an axis the
// message crosses the neighbour's halo layer is 0 (sent in +direction) or n'+1 (sent in -direction); along the other
// axes the neighbour shares this block's index range (Cartesian blocks), L3/ex_sendrecv.f90:12-123.
template <int EX, int EY, int EZ>
__device__ __forceinline__ long long peer_cell(const PeerTable *__restrict__ pt, int D, int i, int j, int k) {
    const int ii = EX > 0 ? 0 : (EX < 0 ? pt->n[D][0] + 1 : i);
    const int jj = EY > 0 ? 0 : (EY < 0 ? pt->n[D][1] + 1 : j);
    const int kk = EZ > 0 ? 0 : (EZ < 0 ? pt->n[D][2] + 1 : k);
    return kk * pt->sz[D] + jj * pt->sy[D] + (ii + OX - 1);
}
#define MGLC_PEER_FACE(D, EX, EY, EZ, a0, a1, a2, a3, a4)                                   \
    if (pm & (1u << (D))) {                                                                   \
        double *P = pt->F[D];                                                                 \
        const long long o = peer_cell<EX, EY, EZ>(pt, D, i, j, k), q = pt->sq[D];             \
        P[(a0) * q + o] = fp[a0]; P[(a1) * q + o] = fp[a1]; P[(a2) * q + o] = fp[a2];         \
        P[(a3) * q + o] = fp[a3]; P[(a4) * q + o] = fp[a4];                                   \
    }
#define MGLC_PEER_EDGE(A, EX, EY, EZ)                                                        \
    if (pm & (1u << (A))) pt->F[A][(A) * pt->sq[A] + peer_cell<EX, EY, EZ>(pt, A, i, j, k)] = fp[A];
// does the CTA whose first cell is (ib, jb, k) contain a cell on a face of the subdomain?  (block-uniform)
// ... and is the neighbour barrier intact?  After a time-out (PeerTable::err) nothing is stored into a neighbour any more:
// the step goes on computing on this subdomain's own stale halos, and the host reports MGLC_E_STATE at every synchronisation.
// Testing the word here -- by the ~1 % of CTAs on a face, after their arithmetic -- keeps two dependent loads off the
// start of every thread (they cost 1 % of the step there: 21.71 instead of 21.52 ms at 768^3 on 2 GPUs).
__device__ __forceinline__ bool peer_cta_on_face(const Geom &g, const PeerTable *__restrict__ pt, int ib, int jb, int k) {
    const bool face = (k == 1) | (k == g.nz) | (jb == 1) | (jb + (int)blockDim.y - 1 >= g.ny) | (ib == 1) | (ib + (int)blockDim.x - 1 >= g.nx);
    return face && !(pt->err && *pt->err);
}
// the outgoing populations of a boundary cell, in the message sets of message_passing_sendrecv()
__device__ __forceinline__ void peer_store_f(const PeerTable *__restrict__ pt, const Geom &g, int i, int j, int k,
                                             const double (&fp)[19]) {
    const unsigned pm = pt->mask;
    const bool xp = i == g.nx, xm = i == 1, yp = j == g.ny, ym = j == 1, zp = k == g.nz, zm = k == 1;
    if (!(xp | xm | yp | ym | zp | zm)) return;
    if (xp) MGLC_PEER_FACE(0, 1, 0, 0, 1, 7, 9, 11, 13)
    if (xm) MGLC_PEER_FACE(1, -1, 0, 0, 2, 8, 10, 12, 14)
    if (yp) MGLC_PEER_FACE(2, 0, 1, 0, 3, 7, 8, 15, 17)
    if (ym) MGLC_PEER_FACE(3, 0, -1, 0, 4, 9, 10, 16, 18)
    if (zp) MGLC_PEER_FACE(4, 0, 0, 1, 5, 11, 12, 15, 16)
    if (zm) MGLC_PEER_FACE(5, 0, 0, -1, 6, 13, 14, 17, 18)
    if (xp && yp) MGLC_PEER_EDGE(7, 1, 1, 0)
    if (xm && yp) MGLC_PEER_EDGE(8, -1, 1, 0)
    if (xp && ym) MGLC_PEER_EDGE(9, 1, -1, 0)
    if (xm && ym) MGLC_PEER_EDGE(10, -1, -1, 0)
    if (xp && zp) MGLC_PEER_EDGE(11, 1, 0, 1)
    if (xm && zp) MGLC_PEER_EDGE(12, -1, 0, 1)
    if (xp && zm) MGLC_PEER_EDGE(13, 1, 0, -1)
    if (xm && zm) MGLC_PEER_EDGE(14, -1, 0, -1)
    if (yp && zp) MGLC_PEER_EDGE(15, 0, 1, 1)
    if (ym && zp) MGLC_PEER_EDGE(16, 0, -1, 1)
    if (yp && zm) MGLC_PEER_EDGE(17, 0, 1, -1)
    if (ym && zm) MGLC_PEER_EDGE(18, 0, -1, -1)
}
// thermal: face d carries the single g population d+1 (B3:1421-1468), no edges
__device__ __forceinline__ void peer_store_g(const PeerTable *__restrict__ pt, const Geom &g, int i, int j, int k,
                                             const double (&gp)[7]) {
    const unsigned pm = pt->mask;
    if (i == g.nx && (pm & 1u)) pt->G[0][1 * pt->sq[0] + peer_cell<1, 0, 0>(pt, 0, i, j, k)] = gp[1];
    if (i == 1 && (pm & 2u)) pt->G[1][2 * pt->sq[1] + peer_cell<-1, 0, 0>(pt, 1, i, j, k)] = gp[2];
    if (j == g.ny && (pm & 4u)) pt->G[2][3 * pt->sq[2] + peer_cell<0, 1, 0>(pt, 2, i, j, k)] = gp[3];
    if (j == 1 && (pm & 8u)) pt->G[3][4 * pt->sq[3] + peer_cell<0, -1, 0>(pt, 3, i, j, k)] = gp[4];
    if (k == g.nz && (pm & 16u)) pt->G[4][5 * pt->sq[4] + peer_cell<0, 0, 1>(pt, 4, i, j, k)] = gp[5];
    if (k == 1 && (pm & 32u)) pt->G[5][6 * pt->sq[5] + peer_cell<0, 0, -1>(pt, 5, i, j, k)] = gp[6];
}

// collision operator selected at compile time: the MRT of L3/collision.f90:20-189 or the BGK alternative of :191-198
template <bool BGK>
__device__ __forceinline__ void collide(const double (&f)[19], double rho, double u, double v, double w, const LbmParams &p,
                                        double (&fp)[19]) {
    if (BGK) d3q19_collide_bgk(f, rho, u, v, w, p.Snu, fp);
    else d3q19_collide(f, rho, u, v, w, p.Snu, p.Sq, fp);
}

// collision(): F (interior) + rho,u,v,w -> Fpost (interior)
template <bool BGK>
__global__ void __launch_bounds__(128) k_collision(Geom g, LbmParams p, const double *__restrict__ F,
                                                   const double *__restrict__ rho, const double *__restrict__ u,
                                                   const double *__restrict__ v, const double *__restrict__ w,
                                                   double *__restrict__ Fpost) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq;
    const long long c = g.idx(0, i, j, k);
    const long long m = g.cell(i, j, k);
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = F[a * sq + c];
    collide<BGK>(f, rho[m], u[m], v[m], w[m], p, fp);
#pragma unroll
    for (int a = 0; a < 19; ++a) Fpost[a * sq + c] = fp[a];
}

// fused: pull (streaming) -> wall bounce-back -> macro -> collide -> store
// PEER: also store the outgoing populations of boundary cells into the neighbours' halos (PeerTable)
template <bool BGK, bool PEER>
__global__ void __launch_bounds__(128, 4) k_fused(Geom g, LbmParams p, const double *__restrict__ Fin,
                                                  double *__restrict__ Fout, const double *__restrict__ rho_lid_in,
                                                  double *__restrict__ rho_lid_out, int i0, int i1, int j0, int j1, int k0,
                                                  const PeerTable *__restrict__ pt) {
    // block (128,1) for full rows, (32,4) for the thin x-slabs of the boundary shell (see launch below)
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y, k = k0 + blockIdx.z;
    if (i > i1 || j > j1) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz;
    const long long c = g.idx(0, i, j, k);
    const WallFlags wf = wall_flags(g, i, j, k);
    double f[19], fp[19];
    MGLC_PULL_ALL();
    MGLC_LID(rho_lid_in);
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    collide<BGK>(f, rho, u, v, w, p, fp);
#pragma unroll
    for (int a = 0; a < 19; ++a) Fout[a * sq + c] = fp[a];
    // Only CTAs that touch a face of the subdomain have anything to send; the test is uniform across the CTA, so the
    // interior (all but ~1 % of the CTAs at 768^3) skips the message code in a handful of instructions (ncu, 2 GPUs:
    // the per-thread face tests cost 5.7 % more instructions than k_fused<..., false>, profiles/r2d_*).
    if (PEER && peer_cta_on_face(g, pt, i0 + (int)(blockIdx.x * blockDim.x), j0 + (int)(blockIdx.y * blockDim.y), k))
        peer_store_f(pt, g, i, j, k, fp);
    // the moving-lid bounce-back of the NEXT step needs this step's rho on the lid plane
    // (L3/bounce_back.f90:77-78 reads rho(i,j,nz) left by the previous macro()); it goes to the other
    // side buffer so that the plane this launch read stays intact for canonicalise()
    if (g.lid && k == g.nz) rho_lid_out[(i - 1) + (long long)g.nx * (j - 1)] = rho;
}

// epilogue of a fused run: pull -> f (pre-collision, as the reference leaves it) and macro fields
__global__ void __launch_bounds__(128) k_stream_macro(Geom g, LbmParams p, const double *__restrict__ Fin,
                                                      double *__restrict__ F, const double *rho_lid_in,
                                                      double *rho_o, double *__restrict__ u_o,
                                                      double *__restrict__ v_o, double *__restrict__ w_o) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz;
    const long long c = g.idx(0, i, j, k);
    const WallFlags wf = wall_flags(g, i, j, k);
    double f[19];
    MGLC_PULL_ALL();
    MGLC_LID(rho_lid_in);     // may alias the top plane of rho_o: read here, written below by the same thread
#pragma unroll
    for (int a = 0; a < 19; ++a) F[a * sq + c] = f[a];
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    const long long m = g.cell(i, j, k);
    rho_o[m] = rho; u_o[m] = u; v_o[m] = v; w_o[m] = w;
}

#ifndef MGLC_HOST_SHIM   // tests/host_shim/lbm_host.cpp runs the kernels above on the CPU and has no <<< >>>
static inline dim3 grid_for(int nxs, int nys, int nzs, int tx) { return dim3((nxs + tx - 1) / tx, nys, nzs); }

int launch_collision(const Geom &g, const LbmParams &p, const double *F, const double *rho, const double *u,
                     const double *v, const double *w, double *Fpost, cudaStream_t s) {
    if (p.bgk) k_collision<true><<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, F, rho, u, v, w, Fpost);
    else k_collision<false><<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, F, rho, u, v, w, Fpost);
    return 1;
}

int launch_fused(const Geom &g, const LbmParams &p, const double *Fin, double *Fout, const double *rho_lid_in,
                 double *rho_lid_out, const int box[6], cudaStream_t s, const PeerTable *pt) {
    const int nxs = box[1] - box[0] + 1, nys = box[3] - box[2] + 1, nzs = box[5] - box[4] + 1;
    if (nxs <= 0 || nys <= 0 || nzs <= 0) return 0;
    // narrow x slabs (the shell next to an x neighbour): keep 128 threads busy by stacking rows
    const dim3 block = nxs <= 8 ? dim3(8, 16) : nxs <= 16 ? dim3(16, 8) : nxs <= 32 ? dim3(32, 4) : dim3(128, 1);
    const dim3 grid((nxs + block.x - 1) / block.x, (nys + block.y - 1) / block.y, nzs);
#define MGLC_LAUNCH_FUSED(B, P) k_fused<B, P><<<grid, block, 0, s>>>(g, p, Fin, Fout, rho_lid_in, rho_lid_out, box[0], box[1], box[2], box[3], box[4], pt)
    if (pt) { if (p.bgk) MGLC_LAUNCH_FUSED(true, true); else MGLC_LAUNCH_FUSED(false, true); }
    else { if (p.bgk) MGLC_LAUNCH_FUSED(true, false); else MGLC_LAUNCH_FUSED(false, false); }
#undef MGLC_LAUNCH_FUSED
    return 1;
}

int launch_stream_macro(const Geom &g, const LbmParams &p, const double *Fin, double *F, const double *rho_lid_in,
                        double *rho, double *u, double *v, double *w, cudaStream_t s) {
    k_stream_macro<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, Fin, F, rho_lid_in, rho, u, v, w);
    return 1;
}
#endif

#include "thermal_kernels.inl"

}  // namespace MGLC_NS
}  // namespace mglc
