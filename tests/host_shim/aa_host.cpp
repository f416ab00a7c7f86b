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
};
}  // namespace

extern "C" {
void *aa_shim_create(int nx, int ny, int nz, double Snu, double Sq, double U0, double rho0, int bgk, int strict_build) {
    Sub *S = new Sub();
    S->g = make_geom(nx, ny, nz);
    for (int f = 0; f < 6; ++f) S->g.wall[f] = 1;
    S->g.lid = 1;
    S->p.Snu = Snu; S->p.Sq = Sq; S->p.U0 = U0; S->p.rho0 = rho0; S->p.bgk = bgk;
    S->layout = AA_NATURAL; S->strict_build = strict_build;
    // NaN everywhere outside the interior: the halo ring must never be read
    S->A.assign((size_t)Q * S->g.sq, __builtin_nan(""));
    const size_t n = (size_t)nx * ny * nz;
    S->rho.assign(n, 0.0); S->u.assign(n, 0.0); S->v.assign(n, 0.0); S->w.assign(n, 0.0);
    S->lid.assign((size_t)nx * ny, __builtin_nan(""));
    return S;
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
    const Geom &g = S->g;
    const LbmParams &p = S->p;
    double *A = S->A.data(), *lid = S->lid.data(), *rho = S->rho.data(), *u = S->u.data(), *v = S->v.data(), *w = S->w.data();
    const unsigned gx = (g.nx + 127) / 128;
    const bool st = S->strict_build, bgk = p.bgk;
    return aa_run(S->layout, nsteps, [&](AaOp op) -> long long {
        switch (op) {
        case AA_OP_LID_PLANE: sweep(((long long)g.nx * g.ny + 255) / 256, 1, 1, 256, [&] { k_aa_lid_plane(g, rho, lid); }); return 1;
        case AA_OP_COLLIDE0:
            if (st) { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_collide0<true>(g, p, A, rho, u, v, w); });
                      else sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_collide0<false>(g, p, A, rho, u, v, w); }); }
            else { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_collide0<true>(g, p, A, rho, u, v, w); });
                   else sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_collide0<false>(g, p, A, rho, u, v, w); }); }
            return 1;
        case AA_OP_ODD:
            if (st) { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_odd<true>(g, p, A, lid); });
                      else sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_odd<false>(g, p, A, lid); }); }
            else { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_odd<true>(g, p, A, lid); });
                   else sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_odd<false>(g, p, A, lid); }); }
            return 1;
        case AA_OP_EVEN:
            if (st) { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_even<true>(g, p, A, lid); });
                      else sweep(gx, g.ny, g.nz, 128, [&] { strict::k_aa_even<false>(g, p, A, lid); }); }
            else { if (bgk) sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_even<true>(g, p, A, lid); });
                   else sweep(gx, g.ny, g.nz, 128, [&] { fast::k_aa_even<false>(g, p, A, lid); }); }
            return 1;
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
    });
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
