// Test-only shim: the product's AA-pattern D3Q19 kernels (mglc_b200/csrc/lbm_aa_kernels.inl, lbm_aa_exact.inl) and their launch
// schedule (aa_run, lbm_aa.cuh) compiled for the HOST and run thread by thread, so the CPU-only suite can check the in-place
// update -- pull / push addressing, the wall rule on both sides of an odd launch, the lid term, resuming from either layout --
// against the oracle without a GPU.  These kernels use no shared memory and no synchronisation and every lattice location is
// touched by exactly one thread, so a sequential sweep over (blockIdx, threadIdx) is an exact emulation.  Strict build:
// bit-identical to the oracle when compiled with -ffp-contract=off; fast build: to rounding.  Never linked into the product.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#define MGLC_HOST_SHIM 1
#define __grid_constant__
#undef __launch_bounds__
#define __launch_bounds__(...)
struct shim_dim3 { unsigned x, y, z; };
static shim_dim3 shim_threadIdx, shim_blockIdx, shim_blockDim;
#define threadIdx shim_threadIdx
#define blockIdx shim_blockIdx
#define blockDim shim_blockDim
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> static inline T __ldg(const T *p) { return *p; }

#define MGLC_NS strict
#define MGLC_STRICT 1
#include "../../mglc_b200/csrc/lbm_aa_kernels.inl"
#undef MGLC_NS
#undef MGLC_STRICT
#undef AA_PARK
#define MGLC_NS fast
#include "../../mglc_b200/csrc/lbm_aa_kernels.inl"
#include "../../mglc_b200/csrc/lbm_aa.cuh"

using namespace mglc;

namespace {
using mglc::strict::AaWalls;
using mglc::strict::aa_walls;
using mglc::strict::d3q19_macro;
#include "../../mglc_b200/csrc/lbm_aa_exact.inl"

template <class K>
void sweep(unsigned gx, unsigned gy, unsigned gz, unsigned bx, K kernel) {
    shim_blockDim = {bx, 1, 1};
    for (unsigned z = 0; z < gz; ++z)
        for (unsigned y = 0; y < gy; ++y)
            for (unsigned x = 0; x < gx; ++x)
                for (unsigned t = 0; t < bx; ++t) {
                    shim_blockIdx = {x, y, z};
                    shim_threadIdx = {t, 0, 0};
                    kernel();
                }
}
struct Sub {
    Geom g;
    LbmParams p;
    int layout, strict_build;
    std::vector<double> A, rho, u, v, w, lid;
    PeerTable pt;          // mask != 0: a block of a decomposed lattice, the neighbours' lattices are the other Subs' vectors
};
// one launch of the schedule on one block; peers = the block's PeerTable or nullptr
long long launch_op(Sub *S, AaOp op) {
    const Geom &g = S->g;
    const LbmParams &p = S->p;
    double *A = S->A.data(), *lid = S->lid.data(), *rho = S->rho.data(), *u = S->u.data(), *v = S->v.data(), *w = S->w.data();
    const unsigned gx = (g.nx + 127) / 128;
    const bool st = S->strict_build, bgk = p.bgk, peer = S->pt.mask != 0;
    const PeerTable &pt = S->pt;
#define SHIM_LAUNCH(K, ...)                                                                                         \
    do {                                                                                                            \
        if (st) { if (bgk) { if (peer) sweep(gx, g.ny, g.nz, 128, [&] { strict::K<true, true>(__VA_ARGS__, pt); });   \
                             else sweep(gx, g.ny, g.nz, 128, [&] { strict::K<true, false>(__VA_ARGS__, pt); }); }     \
                  else { if (peer) sweep(gx, g.ny, g.nz, 128, [&] { strict::K<false, true>(__VA_ARGS__, pt); });      \
                         else sweep(gx, g.ny, g.nz, 128, [&] { strict::K<false, false>(__VA_ARGS__, pt); }); } }      \
        else { if (bgk) { if (peer) sweep(gx, g.ny, g.nz, 128, [&] { fast::K<true, true>(__VA_ARGS__, pt); });        \
                          else sweep(gx, g.ny, g.nz, 128, [&] { fast::K<true, false>(__VA_ARGS__, pt); }); }          \
               else { if (peer) sweep(gx, g.ny, g.nz, 128, [&] { fast::K<false, true>(__VA_ARGS__, pt); });           \
                      else sweep(gx, g.ny, g.nz, 128, [&] { fast::K<false, false>(__VA_ARGS__, pt); }); } }           \
    } while (0)
    switch (op) {
    case AA_OP_LID_PLANE: sweep(((long long)g.nx * g.ny + 255) / 256, 1, 1, 256, [&] { k_aa_lid_plane(g, rho, lid); }); return 1;
    case AA_OP_COLLIDE0: SHIM_LAUNCH(k_aa_collide0, g, p, A, rho, u, v, w); return 1;
    case AA_OP_ODD: SHIM_LAUNCH(k_aa_odd, g, p, A, lid); return 1;
    case AA_OP_EVEN: SHIM_LAUNCH(k_aa_even, g, p, A, lid); return 1;
    case AA_OP_MACRO_POST: sweep(gx, g.ny, g.nz, 128, [&] { k_aa_macro_post(g, p, A, lid, rho, u, v, w); }); return 1;
    case AA_OP_MACRO:       // launch_macro (exact_kernels.cu) = macro() on the natural layout: d3q19_macro per cell
        for (int k = 1; k <= g.nz; ++k)
            for (int j = 1; j <= g.ny; ++j)
                for (int i = 1; i <= g.nx; ++i) {
                    double f[19];
                    for (int a = 0; a < Q; ++a) f[a] = A[g.idx(a, i, j, k)];
                    const long long m = g.cell(i, j, k);
                    d3q19_macro(f, rho[m], u[m], v[m], w[m]);
                }
        return 1;
    }
    return 0;
#undef SHIM_LAUNCH
}
}  // namespace

extern "C" {
// wall[6]: which faces (+x,-x,+y,-y,+z,-z) of the block are walls of the global box; lid: the block touches the lid
void *aa_shim_create_block(int nx, int ny, int nz, const int *wall, int lid, double Snu, double Sq, double U0, double rho0, int bgk,
                           int strict_build) {
    Sub *S = new Sub();
    S->g = make_geom(nx, ny, nz);
    for (int f = 0; f < 6; ++f) S->g.wall[f] = wall[f];
    S->g.lid = lid;
    memset(&S->pt, 0, sizeof S->pt);
    S->p.Snu = Snu; S->p.Sq = Sq; S->p.U0 = U0; S->p.rho0 = rho0; S->p.bgk = bgk;
    S->layout = AA_NATURAL; S->strict_build = strict_build;
    // NaN everywhere outside the interior: the halo ring must never be read
    S->A.assign((size_t)Q * S->g.sq, __builtin_nan(""));
    const size_t n = (size_t)nx * ny * nz;
    S->rho.assign(n, 0.0); S->u.assign(n, 0.0); S->v.assign(n, 0.0); S->w.assign(n, 0.0);
    S->lid.assign((size_t)nx * ny, __builtin_nan(""));
    return S;
}
void *aa_shim_create(int nx, int ny, int nz, double Snu, double Sq, double U0, double rho0, int bgk, int strict_build) {
    const int wall[6] = {1, 1, 1, 1, 1, 1};
    return aa_shim_create_block(nx, ny, nz, wall, 1, Snu, Sq, U0, rho0, bgk, strict_build);
}
// nbr[d], d = 0..18: the block in direction d (faces 0..5, edge populations 7..18) or NULL
void aa_shim_set_peers(void *h, void *const *nbr) {
    Sub *S = (Sub *)h;
    memset(&S->pt, 0, sizeof S->pt);
    for (int d = 0; d < 19; ++d) {
        if (d == 6 || !nbr[d]) continue;
        Sub *N = (Sub *)nbr[d];
        S->pt.mask |= 1u << d;
        S->pt.F[d] = N->A.data(); S->pt.sy[d] = N->g.sy; S->pt.sz[d] = N->g.sz; S->pt.sq[d] = N->g.sq;
        S->pt.n[d][0] = N->g.nx; S->pt.n[d][1] = N->g.ny; S->pt.n[d][2] = N->g.nz;
    }
}
void aa_shim_destroy(void *h) { delete (Sub *)h; }
// f(0:18,nx,ny,nz) + rho,u,v,w in the reference's layout -> NATURAL
void aa_shim_upload(void *h, const double *f, const double *rho, const double *u, const double *v, const double *w) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    for (int k = 1; k <= g.nz; ++k)
        for (int j = 1; j <= g.ny; ++j)
            for (int i = 1; i <= g.nx; ++i)
                for (int a = 0; a < Q; ++a) S->A[g.idx(a, i, j, k)] = f[a + (size_t)Q * g.cell(i, j, k)];
    const size_t b = sizeof(double) * g.nx * g.ny * g.nz;
    memcpy(S->rho.data(), rho, b); memcpy(S->u.data(), u, b); memcpy(S->v.data(), v, b); memcpy(S->w.data(), w, b);
    S->layout = AA_NATURAL;
}
int aa_shim_layout(void *h) { return ((Sub *)h)->layout; }
long long aa_shim_step(void *h, int nsteps) {
    Sub *S = (Sub *)h;
    return aa_run(S->layout, nsteps, [&](AaOp op) -> long long { return launch_op(S, op); });
}
// the blocks of a decomposed lattice advance launch by launch together (mglc_aa_group_step); within one launch the blocks run
// one after the other, in ascending (order = 0) or descending (order = 1) rank order: the results must not depend on it,
// since on the GPU the same launch of neighbouring blocks runs concurrently
long long aa_shim_world_step(void *const *subs, int n, int nsteps, int order) {
    int layout = ((Sub *)subs[0])->layout;
    const long long launches = aa_run(layout, nsteps, [&](AaOp op) -> long long {
        for (int q = 0; q < n; ++q) launch_op((Sub *)subs[order ? n - 1 - q : q], op);
        return 1;
    });
    for (int q = 0; q < n; ++q) ((Sub *)subs[q])->layout = layout;
    return launches;
}
void aa_shim_download_macro(void *h, double *rho, double *u, double *v, double *w) {
    Sub *S = (Sub *)h;
    const size_t b = sizeof(double) * S->g.nx * S->g.ny * S->g.nz;
    memcpy(rho, S->rho.data(), b); memcpy(u, S->u.data(), b); memcpy(v, S->v.data(), b); memcpy(w, S->w.data(), b);
}
// f of the last loop body in the reference's layout, in chunks like mglc_aa_download_f
void aa_shim_download_f(void *h, double *f, long long chunk) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    const long long total = (long long)g.nx * g.ny * g.nz;
    if (S->layout == AA_NATURAL) {
        for (int k = 1; k <= g.nz; ++k)
            for (int j = 1; j <= g.ny; ++j)
                for (int i = 1; i <= g.nx; ++i)
                    for (int a = 0; a < Q; ++a) f[a + (size_t)Q * g.cell(i, j, k)] = S->A[g.idx(a, i, j, k)];
        return;
    }
    for (long long c0 = 0; c0 < total; c0 += chunk) {
        const long long nc = total - c0 < chunk ? total - c0 : chunk;
        sweep((nc + 127) / 128, 1, 1, 128, [&] { k_aa_gather_f(g, S->p, S->A.data(), S->lid.data(), c0, nc, f + c0 * Q); });
    }
}
}
