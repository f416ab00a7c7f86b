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
#include "d3q19_mrt.inl"
#include "d3q19_thermal.inl"

namespace mglc {
namespace MGLC_NS {

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
template <bool BGK>
__global__ void __launch_bounds__(128, 4) k_fused(Geom g, LbmParams p, const double *__restrict__ Fin,
                                                  double *__restrict__ Fout, const double *__restrict__ rho_lid_in,
                                                  double *__restrict__ rho_lid_out, int i0, int i1, int j0, int j1, int k0) {
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

static inline dim3 grid_for(int nxs, int nys, int nzs, int tx) { return dim3((nxs + tx - 1) / tx, nys, nzs); }

int launch_collision(const Geom &g, const LbmParams &p, const double *F, const double *rho, const double *u,
                     const double *v, const double *w, double *Fpost, cudaStream_t s) {
    if (p.bgk) k_collision<true><<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, F, rho, u, v, w, Fpost);
    else k_collision<false><<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, F, rho, u, v, w, Fpost);
    return 1;
}

int launch_fused(const Geom &g, const LbmParams &p, const double *Fin, double *Fout, const double *rho_lid_in,
                 double *rho_lid_out, const int box[6], cudaStream_t s) {
    const int nxs = box[1] - box[0] + 1, nys = box[3] - box[2] + 1, nzs = box[5] - box[4] + 1;
    if (nxs <= 0 || nys <= 0 || nzs <= 0) return 0;
    const dim3 block = nxs <= 32 ? dim3(32, 4) : dim3(128, 1);
    const dim3 grid((nxs + block.x - 1) / block.x, (nys + block.y - 1) / block.y, nzs);
    if (p.bgk) k_fused<true><<<grid, block, 0, s>>>(g, p, Fin, Fout, rho_lid_in, rho_lid_out, box[0], box[1], box[2], box[3], box[4]);
    else k_fused<false><<<grid, block, 0, s>>>(g, p, Fin, Fout, rho_lid_in, rho_lid_out, box[0], box[1], box[2], box[3], box[4]);
    return 1;
}

int launch_stream_macro(const Geom &g, const LbmParams &p, const double *Fin, double *F, const double *rho_lid_in,
                        double *rho, double *u, double *v, double *w, cudaStream_t s) {
    k_stream_macro<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, Fin, F, rho_lid_in, rho, u, v, w);
    return 1;
}

#include "thermal_kernels.inl"

}  // namespace MGLC_NS
}  // namespace mglc
