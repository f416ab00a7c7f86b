// Test-only shim: the product's 2-D lid-driven cavity kernels (mglc_b200/csrc/lid2d_kernels.inl: k_l2_collision, k_l2_fused,
// k_l2_stream_macro, all four arithmetics; lid2d_exact.inl: k_l2_initial, k_l2_streaming, k_l2_bounceback, k_l2_macro)
// compiled for the HOST and run thread by thread, so the CPU-only suite can check
// the pull addressing, the wall rule and the lid term (incl. the two top corners) against the oracle without a GPU.  No shared
// memory, no synchronisation: a sequential sweep over (blockIdx, threadIdx) is an exact emulation.  Never linked into the product.
#include <cuda_runtime.h>

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
#include "../../mglc_b200/csrc/lid2d_kernels.inl"
#undef MGLC_NS
#undef MGLC_STRICT
#define MGLC_NS fast
#include "../../mglc_b200/csrc/lid2d_kernels.inl"

using namespace mglc;

namespace {
#include "../../mglc_b200/csrc/lid2d_exact.inl"

template <class K>
void sweep(const Geom2 &g, K kernel) {
    shim_blockDim = {128, 1, 1};
    for (unsigned by = 0; by < (unsigned)g.ny; ++by)
        for (unsigned bx = 0; bx < (unsigned)((g.nx + 127) / 128); ++bx)
            for (unsigned tx = 0; tx < 128; ++tx) {
                shim_blockIdx = {bx, by, 0};
                shim_threadIdx = {tx, 0, 0};
                kernel();
            }
}
void to_soa(const Geom2 &g, const double *aos, std::vector<double> &P) {
    P.assign((size_t)9 * g.sq, 0.0);
    for (int j = 0; j <= g.ny + 1; ++j)
        for (int i = 0; i <= g.nx + 1; ++i)
            for (int a = 0; a < 9; ++a) P[g.idx(a, i, j)] = aos[a + (size_t)9 * (i + (size_t)(g.nx + 2) * j)];
}
void to_aos(const Geom2 &g, const std::vector<double> &P, double *aos) {
    for (int j = 0; j <= g.ny + 1; ++j)
        for (int i = 0; i <= g.nx + 1; ++i)
            for (int a = 0; a < 9; ++a) aos[a + (size_t)9 * (i + (size_t)(g.nx + 2) * j)] = P[g.idx(a, i, j)];
}
}  // namespace

extern "C" {
// mode 0: k_l2_fused         f_post (halo'd) + lid_in -> f_post_out (halo'd, interior written), lid_out
// mode 1: k_l2_stream_macro  f_post + lid_in -> f_out (halo'd array, interior written), fields3 = rho,u,v
// mode 2: k_l2_collision     fin holds f (halo'd array, interior used), fields3 = rho,u,v in -> f_post_out
// modes 3..6: the per-subroutine kernels initial / streaming / bounceback / macro (see below); variant 0 = L2C, 1 = L2F, 2 = L2I, 3 = L2C with model = SRT
int shim_l2d(int mode, int strict_build, int variant, int nx, int ny, const int *wall, double Snu, double Sq, double U0, double rho0,
             const double *fin, const double *lid_in, double *fout, double *lid_out, double *fields3) {
    Geom2 g = make_geom2(nx, ny);
    for (int q = 0; q < 4; ++q) g.wall[q] = wall[q];
    const L2Params p{Snu, Sq, U0, rho0};
    std::vector<double> Fi, Fo((size_t)9 * g.sq, 0.0);
    to_soa(g, fin, Fi);
    const size_t n = (size_t)nx * ny;
    double *rho = fields3, *u = fields3 + n, *v = fields3 + 2 * n;
    const double *fi = Fi.data();
    double *fo = Fo.data();
#define L2_BY_VARIANT(NS, KERNEL, ...)                                              \
    {                                                                               \
        if (variant == 0) sweep(g, [&] { NS::KERNEL<0>(__VA_ARGS__); });            \
        else if (variant == 1) sweep(g, [&] { NS::KERNEL<1>(__VA_ARGS__); });       \
        else if (variant == 2) sweep(g, [&] { NS::KERNEL<2>(__VA_ARGS__); });       \
        else sweep(g, [&] { NS::KERNEL<3>(__VA_ARGS__); });                         \
    }
    const bool inc = variant == 2;
    if (mode == 0) {
        if (strict_build) L2_BY_VARIANT(strict, k_l2_fused, g, p, fi, fo, lid_in, lid_out)
        else L2_BY_VARIANT(fast, k_l2_fused, g, p, fi, fo, lid_in, lid_out)
    } else if (mode == 1) {
        if (inc) sweep(g, [&] { strict::k_l2_stream_macro<true>(g, p, fi, fo, lid_in, rho, u, v); });
        else sweep(g, [&] { strict::k_l2_stream_macro<false>(g, p, fi, fo, lid_in, rho, u, v); });
    } else if (mode == 2) {
        if (strict_build) L2_BY_VARIANT(strict, k_l2_collision, g, p, fi, rho, u, v, fo)
        else L2_BY_VARIANT(fast, k_l2_collision, g, p, fi, rho, u, v, fo)
    } else if (mode == 3) {          // k_l2_initial: -> f (interior of fout), rho, u, v; up, vp are zero-filled scratch
        std::vector<double> up(n), vp(n);
        if (inc) sweep(g, [&] { k_l2_initial<true>(g, p, g.wall[2], fo, rho, u, v, up.data(), vp.data()); });
        else sweep(g, [&] { k_l2_initial<false>(g, p, g.wall[2], fo, rho, u, v, up.data(), vp.data()); });
    } else if (mode == 4) {          // k_l2_streaming: fin = f_post -> f
        sweep(g, [&] { k_l2_streaming(g, fi, fo); });
    } else if (mode == 5) {          // k_l2_bounceback: fin = f_post, fout holds f on entry (in place), fields3[0] = rho
        to_soa(g, fout, Fo);
        fo = Fo.data();
        const int cells = 2 * nx + 2 * (ny - 2 > 0 ? ny - 2 : 0);
        shim_blockDim = {128, 1, 1};
        for (unsigned bx = 0; bx < (unsigned)((cells + 127) / 128); ++bx)
            for (unsigned tx = 0; tx < 128; ++tx) {
                shim_blockIdx = {bx, 0, 0};
                shim_threadIdx = {tx, 0, 0};
                if (inc) k_l2_bounceback<true>(g, p, fi, rho, fo); else k_l2_bounceback<false>(g, p, fi, rho, fo);
            }
    } else if (mode == 6) {          // k_l2_macro: fin = f -> rho, u, v
        if (inc) sweep(g, [&] { k_l2_macro<true>(g, fi, rho, u, v); });
        else sweep(g, [&] { k_l2_macro<false>(g, fi, rho, u, v); });
    } else return -1;
    to_aos(g, Fo, fout);
    return 0;
}
}
