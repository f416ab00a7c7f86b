// Test-only shim: the product's 3-D D3Q19 kernels (mglc_b200/csrc/lbm_kernels.inl: k_collision, k_fused with and without the
// direct halo stores into the neighbours' lattices, k_stream_macro; thermal_kernels.inl: k_th_fused) compiled for the HOST and run
// thread by thread, so the CPU-only suite can check the pull addressing, the unified wall rule, the lid term and -- on P emulated
// subdomains -- the PeerTable stores (faces, edges, the thermal g population) against the oracle without a GPU.  The kernels use
// no shared memory and no synchronisation; a launch only writes the OTHER lattice, so sweeping the subdomains one after the
// other is an exact emulation of P concurrent launches.  Never linked into the product.
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
#include "../../mglc_b200/csrc/lbm_kernels.inl"
#undef MGLC_NS
#undef MGLC_STRICT
#define MGLC_NS fast
#include "../../mglc_b200/csrc/lbm_kernels.inl"

using namespace mglc;

namespace {
template <class K>
void sweep(unsigned gx, unsigned gy, unsigned gz, K kernel) {
    shim_blockDim = {128, 1, 1};
    for (unsigned z = 0; z < gz; ++z)
        for (unsigned y = 0; y < gy; ++y)
            for (unsigned x = 0; x < gx; ++x)
                for (unsigned t = 0; t < 128; ++t) {
                    shim_blockIdx = {x, y, z};
                    shim_threadIdx = {t, 0, 0};
                    kernel();
                }
}
struct Sub {
    Geom g;
    LbmParams p;
    ThermalParams tp;
    int strict_build;
    std::vector<double> F[2], G[2], Fc[2], lid[2];      // two lattices each, like the product's ping-pong buffers
    PeerTable pt[2];                                    // pt[b]: the table of a launch that writes lattice b
};
}  // namespace

extern "C" {
// wall[6], lid as Geom; par = Snu, Sq, U0, rho0, bgk ; tpar (thermal, may be NULL) = Snu, Sq, Qd, Qnu, paraA, gBeta, Tref, omegaRot,
// Thot, Tcold, wallT[6], bcT[6]
void *lbm_shim_create(int nx, int ny, int nz, const int *wall, int lid, const double *par, const double *tpar, int strict_build) {
    Sub *S = new Sub();
    S->g = make_geom(nx, ny, nz);
    for (int f = 0; f < 6; ++f) S->g.wall[f] = wall[f];
    S->g.lid = lid;
    S->p.Snu = par[0]; S->p.Sq = par[1]; S->p.U0 = par[2]; S->p.rho0 = par[3]; S->p.bgk = (int)par[4];
    S->strict_build = strict_build;
    memset(&S->tp, 0, sizeof S->tp);
    if (tpar) {
        ThermalParams &t = S->tp;
        t.Snu = tpar[0]; t.Sq = tpar[1]; t.Qd = tpar[2]; t.Qnu = tpar[3]; t.paraA = tpar[4]; t.gBeta = tpar[5]; t.Tref = tpar[6];
        t.omegaRot = tpar[7]; t.Thot = tpar[8]; t.Tcold = tpar[9];
        for (int f = 0; f < 6; ++f) { t.wallT[f] = tpar[10 + f]; t.bcT[f] = (int)tpar[16 + f]; }
    }
    const double nan = __builtin_nan("");
    const size_t n = (size_t)nx * ny * nz;
    for (int b = 0; b < 2; ++b) {
        S->F[b].assign((size_t)Q * S->g.sq, nan);
        S->G[b].assign((size_t)QT * S->g.sq, nan);
        S->Fc[b].assign(3 * n, nan);
        S->lid[b].assign((size_t)nx * ny, nan);
        memset(&S->pt[b], 0, sizeof(PeerTable));
    }
    return S;
}
void lbm_shim_destroy(void *h) { delete (Sub *)h; }
// lattice `b` of f (nq = 19) or g (nq = 7) <-> the reference layout WITH halos, (0:nq-1, 0:nx+1, 0:ny+1, 0:nz+1)
void lbm_shim_put(void *h, int nq, int b, const double *aos) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    std::vector<double> &L = nq == Q ? S->F[b] : S->G[b];
    for (int k = 0; k <= g.nz + 1; ++k)
        for (int j = 0; j <= g.ny + 1; ++j)
            for (int i = 0; i <= g.nx + 1; ++i)
                for (int a = 0; a < nq; ++a) L[g.idx(a, i, j, k)] = aos[a + (size_t)nq * (i + (size_t)(g.nx + 2) * (j + (size_t)(g.ny + 2) * k))];
}
void lbm_shim_get(void *h, int nq, int b, double *aos) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    const std::vector<double> &L = nq == Q ? S->F[b] : S->G[b];
    for (int k = 0; k <= g.nz + 1; ++k)
        for (int j = 0; j <= g.ny + 1; ++j)
            for (int i = 0; i <= g.nx + 1; ++i)
                for (int a = 0; a < nq; ++a) aos[a + (size_t)nq * (i + (size_t)(g.nx + 2) * (j + (size_t)(g.ny + 2) * k))] = L[g.idx(a, i, j, k)];
}
void lbm_shim_put_lid(void *h, int b, const double *plane) { Sub *S = (Sub *)h; memcpy(S->lid[b].data(), plane, sizeof(double) * S->g.nx * S->g.ny); }
void lbm_shim_get_lid(void *h, int b, double *plane) { Sub *S = (Sub *)h; memcpy(plane, S->lid[b].data(), sizeof(double) * S->g.nx * S->g.ny); }
void lbm_shim_put_force(void *h, int b, const double *fc3) { Sub *S = (Sub *)h; memcpy(S->Fc[b].data(), fc3, sizeof(double) * 3 * S->g.nx * S->g.ny * S->g.nz); }
void lbm_shim_get_force(void *h, int b, double *fc3) { Sub *S = (Sub *)h; memcpy(fc3, S->Fc[b].data(), sizeof(double) * 3 * S->g.nx * S->g.ny * S->g.nz); }
// the neighbour in message direction d (0..5 faces, 7..18 edges) of this subdomain, for launches that write lattice b
void lbm_shim_set_peer(void *h, int b, int d, void *neighbour) {
    Sub *S = (Sub *)h, *N = (Sub *)neighbour;
    PeerTable &t = S->pt[b];
    t.mask |= 1u << d;
    t.F[d] = N->F[b].data();
    if (d < 6) t.G[d] = N->G[b].data();
    t.sy[d] = N->g.sy; t.sz[d] = N->g.sz; t.sq[d] = N->g.sq;
    t.n[d][0] = N->g.nx; t.n[d][1] = N->g.ny; t.n[d][2] = N->g.nz;
}
// k_fused: lattice `in` -> lattice in^1 over the whole block, with (peers != 0) or without the direct halo stores
void lbm_shim_fused(void *h, int in, int peers) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    const int out = in ^ 1;
    const double *Fin = S->F[in].data(), *li = S->lid[in].data();
    double *Fout = S->F[out].data(), *lo = S->lid[out].data();
    const PeerTable *pt = &S->pt[out];
    const unsigned gx = (g.nx + 127) / 128;
#define RUN(NS, B, P) sweep(gx, g.ny, g.nz, [&] { NS::k_fused<B, P>(g, S->p, Fin, Fout, li, lo, 1, g.nx, 1, g.ny, 1, pt); })
    if (S->strict_build) { if (S->p.bgk) { if (peers) RUN(strict, true, true); else RUN(strict, true, false); } else { if (peers) RUN(strict, false, true); else RUN(strict, false, false); } }
    else { if (S->p.bgk) { if (peers) RUN(fast, true, true); else RUN(fast, true, false); } else { if (peers) RUN(fast, false, true); else RUN(fast, false, false); } }
#undef RUN
}
// k_th_fused: f, g, carried force: lattices / force buffer `in` -> in^1
void lbm_shim_th_fused(void *h, int in, int peers) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    const int out = in ^ 1;
    const PeerTable *pt = &S->pt[out];
    const unsigned gx = (g.nx + 127) / 128;
#define RUN(NS, P) sweep(gx, g.ny, g.nz, [&] { NS::k_th_fused<P>(g, S->tp, S->F[in].data(), S->F[out].data(), S->G[in].data(), S->G[out].data(), \
                                                                  S->Fc[in].data(), S->Fc[out].data(), 1, g.nx, 1, g.ny, 1, pt); })
    if (S->strict_build) { if (peers) RUN(strict, true); else RUN(strict, false); }
    else { if (peers) RUN(fast, true); else RUN(fast, false); }
#undef RUN
}
// k_stream_macro: lattice `in` -> f (reference layout without halos) and rho,u,v,w
void lbm_shim_stream_macro(void *h, int in, double *f, double *rho, double *u, double *v, double *w) {
    Sub *S = (Sub *)h;
    const Geom &g = S->g;
    std::vector<double> Fo((size_t)Q * g.sq, 0.0);
    sweep((g.nx + 127) / 128, g.ny, g.nz, [&] { strict::k_stream_macro(g, S->p, S->F[in].data(), Fo.data(), S->lid[in].data(), rho, u, v, w); });
    for (int k = 1; k <= g.nz; ++k)
        for (int j = 1; j <= g.ny; ++j)
            for (int i = 1; i <= g.nx; ++i)
                for (int a = 0; a < Q; ++a) f[a + (size_t)Q * g.cell(i, j, k)] = Fo[g.idx(a, i, j, k)];
}
}
