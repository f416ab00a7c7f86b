// Test-only shim: compiles the product's 2-D thermal kernels (mglc_b200/csrc/thermal2d_kernels.inl + d2q9_thermal.inl: the fused
// pull + macro + collide kernel, the epilogue and the two collision kernels) for the HOST and runs them thread by thread, so
// the CPU-only suite can check the kernels' indexing, wall rule and arithmetic against the oracle without a GPU.  These
// kernels use no shared memory and no synchronisation, so a sequential sweep over (blockIdx, threadIdx) is an exact emulation
// -- including the in-place update of Fy.  Two copies of the same source: MGLC_STRICT (must be bit-identical to the oracle
// when built with -ffp-contract=off) and the throughput form (must agree to rounding).  Never linked into the product.
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
#include "../../mglc_b200/csrc/thermal2d_kernels.inl"
#undef MGLC_NS
#undef MGLC_STRICT
#define MGLC_NS fast
#include "../../mglc_b200/csrc/thermal2d_kernels.inl"

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
// reference layout (population fastest, with a one-cell halo ring) <-> the device's padded SoA rows
void to_soa(const Geom2 &g, int nq, const double *aos, std::vector<double> &P) {
    P.assign((size_t)nq * g.sq, 0.0);
    for (int j = 0; j <= g.ny + 1; ++j)
        for (int i = 0; i <= g.nx + 1; ++i)
            for (int a = 0; a < nq; ++a) P[g.idx(a, i, j)] = aos[a + (size_t)nq * (i + (size_t)(g.nx + 2) * j)];
}
void to_aos(const Geom2 &g, int nq, const std::vector<double> &P, double *aos) {
    for (int j = 0; j <= g.ny + 1; ++j)
        for (int i = 0; i <= g.nx + 1; ++i)
            for (int a = 0; a < nq; ++a) aos[a + (size_t)nq * (i + (size_t)(g.nx + 2) * j)] = P[g.idx(a, i, j)];
}
}  // namespace

extern "C" {
// par = Snu, Sq, Qd, Qnu, paraA, gBeta, Tref, rho0, Thot, Tcold, perx, variant, Uwall[8], cornersT, start[2], total[2] (25 doubles);
// wallT[4], bcT[4], wall[4] as in T2Params / Geom2.  fields4[0] = rho: read at the wall cells when walls move (mode 0 rewrites them).
// mode 0: k_t2_fused         f_post, g_post (halo'd, in) -> f_post_out, g_post_out (halo'd, interior written), Fy in place
// mode 1: k_t2_stream_macro  f_post, g_post -> f_out, g_out (halo'd arrays, interior written), rho,u,v,T (nx*ny each, in fields4)
// mode 2: k_t2_collision + k_t2_collisionT: f_post/g_post hold f/g (halo'd arrays, interior used), fields4 = rho,u,v,T in;
//         -> f_post_out, g_post_out, Fy
int shim_t2d(int mode, int strict_build, int nx, int ny, const int *wall, const double *par, const double *wallT, const int *bcT,
             const double *fin, const double *gin, double *fout, double *gout, double *Fy, double *fields4) {
    Geom2 g = make_geom2(nx, ny);
    for (int q = 0; q < 4; ++q) g.wall[q] = wall[q];
    T2Params p{};
    p.Snu = par[0]; p.Sq = par[1]; p.Qd = par[2]; p.Qnu = par[3]; p.paraA = par[4]; p.gBeta = par[5]; p.Tref = par[6]; p.rho0 = par[7];
    p.Thot = par[8]; p.Tcold = par[9]; p.perx = (int)par[10]; p.variant = (int)par[11];
    p.moving = 0;
    for (int q = 0; q < 8; ++q) { p.Uwall[q] = par[12 + q]; p.moving |= p.Uwall[q] != 0.0; }
    p.cornersT = (int)par[20]; p.start[0] = (int)par[21]; p.start[1] = (int)par[22]; p.total[0] = (int)par[23]; p.total[1] = (int)par[24];
    for (int q = 0; q < 4; ++q) { p.wallT[q] = wallT[q]; p.bcT[q] = bcT[q]; }
    std::vector<double> Fi, Gi, Fo((size_t)9 * g.sq, 0.0), Go((size_t)5 * g.sq, 0.0);
    to_soa(g, 9, fin, Fi); to_soa(g, 5, gin, Gi);
    const size_t n = (size_t)nx * ny;
    double *rho = fields4, *u = fields4 + n, *v = fields4 + 2 * n, *T = fields4 + 3 * n;
    std::vector<double> Fx(n, -1.0);
    if (mode == 0) {
        if (strict_build) sweep(g, [&] { strict::k_t2_fused(g, p, Fi.data(), Fo.data(), Gi.data(), Go.data(), Fy, rho); });
        else sweep(g, [&] { fast::k_t2_fused(g, p, Fi.data(), Fo.data(), Gi.data(), Go.data(), Fy, rho); });
    } else if (mode == 1) {
        if (strict_build) sweep(g, [&] { strict::k_t2_stream_macro(g, p, Fi.data(), Fo.data(), Gi.data(), Go.data(), Fy, rho, u, v, T); });
        else sweep(g, [&] { fast::k_t2_stream_macro(g, p, Fi.data(), Fo.data(), Gi.data(), Go.data(), Fy, rho, u, v, T); });
    } else if (mode == 2) {
        if (strict_build) {
            sweep(g, [&] { strict::k_t2_collision(g, p, Fi.data(), rho, u, v, T, Fo.data(), Fx.data(), Fy); });
            sweep(g, [&] { strict::k_t2_collisionT(g, p, Gi.data(), u, v, T, Go.data()); });
        } else {
            sweep(g, [&] { fast::k_t2_collision(g, p, Fi.data(), rho, u, v, T, Fo.data(), Fx.data(), Fy); });
            sweep(g, [&] { fast::k_t2_collisionT(g, p, Gi.data(), u, v, T, Go.data()); });
        }
        for (double x : Fx) if (x != 0.0) return -2;
    } else return -1;
    to_aos(g, 9, Fo, fout); to_aos(g, 5, Go, gout);
    return 0;
}
}

// ---------------------------------------------------------------------------------------------------------------------
// The copy-type kernels (mglc_b200/csrc/thermal2d_exact.inl: initial, streaming(T), bounceback(T), macro(T), halo pack / unpack,
// layout transposes) on the CPU: one emulated subdomain with the device's arrays; the test drives P of them like the host code
// drives P subdomains and moves the packed message buffers between them.
namespace {
#include "../../mglc_b200/csrc/thermal2d_exact.inl"

template <class K>
void sweep_grid(unsigned gx, unsigned gy, unsigned bx, K kernel) {
    shim_blockDim = {bx, 1, 1};
    for (unsigned y = 0; y < gy; ++y)
        for (unsigned x = 0; x < gx; ++x)
            for (unsigned t = 0; t < bx; ++t) {
                shim_blockIdx = {x, y, 0};
                shim_threadIdx = {t, 0, 0};
                kernel();
            }
}
struct Sub {
    Geom2 g;
    T2Params p;
    std::vector<double> F, G, P, Q, fld[9], stage;      // fld: rho,u,v,T,up,vp,Tp,Fx,Fy
};
}  // namespace

extern "C" {
void *shim_sub_create(int nx, int ny, const int *wall, const double *par, const double *wallT, const int *bcT) {
    Sub *S = new Sub();
    S->g = make_geom2(nx, ny);
    for (int q = 0; q < 4; ++q) S->g.wall[q] = wall[q];
    T2Params &p = S->p;
    p = T2Params{};
    p.Snu = par[0]; p.Sq = par[1]; p.Qd = par[2]; p.Qnu = par[3]; p.paraA = par[4]; p.gBeta = par[5]; p.Tref = par[6]; p.rho0 = par[7];
    p.Thot = par[8]; p.Tcold = par[9]; p.perx = (int)par[10]; p.variant = (int)par[11];
    p.moving = 0;
    for (int q = 0; q < 8; ++q) { p.Uwall[q] = par[12 + q]; p.moving |= p.Uwall[q] != 0.0; }
    p.cornersT = (int)par[20]; p.start[0] = (int)par[21]; p.start[1] = (int)par[22]; p.total[0] = (int)par[23]; p.total[1] = (int)par[24];
    for (int q = 0; q < 4; ++q) { p.wallT[q] = wallT[q]; p.bcT[q] = bcT[q]; }
    S->F.assign((size_t)9 * S->g.sq, 0.0); S->P.assign((size_t)9 * S->g.sq, 0.0);
    S->G.assign((size_t)5 * S->g.sq, 0.0); S->Q.assign((size_t)5 * S->g.sq, 0.0);
    for (auto &f : S->fld) f.assign((size_t)nx * ny, 0.0);
    S->stage.assign((size_t)9 * (nx + 2) * (ny + 2), 0.0);
    return S;
}
void shim_sub_destroy(void *h) { delete (Sub *)h; }
// which: 0 f, 1 f_post, 2 g, 3 g_post (through the transposing kernels); 4.. = fields rho,u,v,T,up,vp,Tp,Fx,Fy (plain copies)
int shim_sub_put(void *h, int which, const double *host) {
    Sub *S = (Sub *)h;
    const Geom2 &g = S->g;
    if (which >= 4) { memcpy(S->fld[which - 4].data(), host, sizeof(double) * g.nx * g.ny); return 0; }
    const int nq = which < 2 ? 9 : 5, halo = which & 1;
    const size_t cells = halo ? (size_t)(g.nx + 2) * (g.ny + 2) : (size_t)g.nx * g.ny;
    memcpy(S->stage.data(), host, sizeof(double) * nq * cells);
    double *dev = which == 0 ? S->F.data() : which == 1 ? S->P.data() : which == 2 ? S->G.data() : S->Q.data();
    sweep_grid((g.nx + 2 * halo + 127) / 128, g.ny + 2 * halo, 128, [&] { k_t2_aos_to_soa(g, nq, S->stage.data(), dev, halo); });
    return 0;
}
int shim_sub_get(void *h, int which, double *host) {
    Sub *S = (Sub *)h;
    const Geom2 &g = S->g;
    if (which >= 4) { memcpy(host, S->fld[which - 4].data(), sizeof(double) * g.nx * g.ny); return 0; }
    const int nq = which < 2 ? 9 : 5, halo = which & 1;
    const size_t cells = halo ? (size_t)(g.nx + 2) * (g.ny + 2) : (size_t)g.nx * g.ny;
    const double *dev = which == 0 ? S->F.data() : which == 1 ? S->P.data() : which == 2 ? S->G.data() : S->Q.data();
    sweep_grid((g.nx + 2 * halo + 127) / 128, g.ny + 2 * halo, 128, [&] { k_t2_soa_to_aos(g, nq, dev, S->stage.data(), halo); });
    memcpy(host, S->stage.data(), sizeof(double) * nq * cells);
    return 0;
}
// op: 0 initial(profile, start, total) 1 streaming 2 streamingT 3 bounceback 4 bouncebackT 5 macro 6 macroT
int shim_sub_op(void *h, int op, int a0, int a1, int a2) {
    Sub *S = (Sub *)h;
    const Geom2 &g = S->g;
    double **f = nullptr; (void)f;
    double *rho = S->fld[0].data(), *u = S->fld[1].data(), *v = S->fld[2].data(), *T = S->fld[3].data(), *up = S->fld[4].data(),
           *vp = S->fld[5].data(), *Tp = S->fld[6].data(), *Fx = S->fld[7].data(), *Fy = S->fld[8].data();
    const unsigned gx = (g.nx + 127) / 128, ring = (2 * g.nx + 2 * (g.ny > 2 ? g.ny - 2 : 0) + 127) / 128;
    switch (op) {
    case 0: sweep_grid(gx, g.ny, 128, [&] { k_t2_initial(g, S->p, a0, a1, a2, S->F.data(), S->G.data(), rho, u, v, T, up, vp, Tp); });
            std::fill(S->P.begin(), S->P.end(), 0.0); std::fill(S->Q.begin(), S->Q.end(), 0.0); break;
    case 1: sweep_grid(gx, g.ny, 128, [&] { k_t2_streaming(g, 9, S->P.data(), S->F.data()); }); break;
    case 2: sweep_grid(gx, g.ny, 128, [&] { k_t2_streaming(g, 5, S->Q.data(), S->G.data()); }); break;
    case 3: sweep_grid(ring, 1, 128, [&] { k_t2_bounceback(g, S->p, S->P.data(), S->F.data(), rho); }); break;
    case 4: sweep_grid(ring, 1, 128, [&] { k_t2_bouncebackT(g, S->p, S->Q.data(), S->G.data()); }); break;
    case 5: sweep_grid(gx, g.ny, 128, [&] { k_t2_macro(g, S->F.data(), Fx, Fy, rho, u, v); }); break;
    case 6: sweep_grid(gx, g.ny, 128, [&] { k_t2_macroT(g, S->G.data(), T); }); break;
    default: return -1;
    }
    return 0;
}
// message dir 0..11 (thermal2d.cu: 0..7 f faces + corners, 8..11 g faces): n1 = extent along the face, npop = populations carried
int shim_sub_pack(void *h, int dir, int n1, int npop, double *buf) {
    Sub *S = (Sub *)h;
    sweep_grid((n1 * npop + 127) / 128, 1, 128, [&] { k_t2_pack(S->g, dir >= 8 ? S->Q.data() : S->P.data(), dir, n1, npop, buf); });
    return 0;
}
int shim_sub_unpack(void *h, int dir, int n1, int npop, const double *buf) {
    Sub *S = (Sub *)h;
    sweep_grid((n1 * npop + 127) / 128, 1, 128, [&] { k_t2_unpack(S->g, dir >= 8 ? S->Q.data() : S->P.data(), dir, n1, npop, buf); });
    return 0;
}
}
