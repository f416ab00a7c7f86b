// Test-only shim: the product's 2-D lid-driven cavity kernels (mglc_b200/csrc/lid2d_kernels.inl: k_l2_collision, k_l2_fused,
// k_l2_stream_macro, both programs' roundings) compiled for the HOST and run thread by thread, so the CPU-only suite can check
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
    if (mode == 0) {
        if (strict_build) { if (variant) sweep(g, [&] { strict::k_l2_fused<1>(g, p, fi, fo, lid_in, lid_out); }); else sweep(g, [&] { strict::k_l2_fused<0>(g, p, fi, fo, lid_in, lid_out); }); }
        else { if (variant) sweep(g, [&] { fast::k_l2_fused<1>(g, p, fi, fo, lid_in, lid_out); }); else sweep(g, [&] { fast::k_l2_fused<0>(g, p, fi, fo, lid_in, lid_out); }); }
    } else if (mode == 1) {
        sweep(g, [&] { strict::k_l2_stream_macro(g, p, fi, fo, lid_in, rho, u, v); });
    } else if (mode == 2) {
        if (strict_build) { if (variant) sweep(g, [&] { strict::k_l2_collision<1>(g, p, fi, rho, u, v, fo); }); else sweep(g, [&] { strict::k_l2_collision<0>(g, p, fi, rho, u, v, fo); }); }
        else { if (variant) sweep(g, [&] { fast::k_l2_collision<1>(g, p, fi, rho, u, v, fo); }); else sweep(g, [&] { fast::k_l2_collision<0>(g, p, fi, rho, u, v, fo); }); }
    } else return -1;
    to_aos(g, Fo, fout);
    return 0;
}
}
