// lbm_aa.cu -- the D3Q19 lid-driven cavity loop (L3/main.f90:85-103) on ONE lattice (AA-pattern storage)
// behind the mglc_aa_* entry points of mglc.h: the same arithmetic and the same results as mglc_lbm_step (bit-identical in
// strict mode), but 19 x 8 B per cell of lattice instead of 2 x 19 x 8 B, for the largest lattices one GPU can hold.
// One subdomain only (all six faces are walls); the ping-pong path (api.cu) is the one that decomposes.
// This file holds the strict build of the kernels (-fmad=false), the order-preserving kernels and the host side.
#define MGLC_NS strict
#define MGLC_STRICT 1
#include "lbm_aa_kernels.inl"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "halo.cuh"
#include "lbm_aa.cuh"

using namespace mglc;

namespace {
using mglc::strict::AaWalls;
using mglc::strict::aa_walls;
using mglc::strict::d3q19_macro;
#include "lbm_aa_exact.inl"
}  // namespace

struct mglc_aa {
    mglc_aa_desc d;
    Geom g;
    LbmParams p;
    double *A;                       // the lattice
    int layout;                      // AA_NATURAL: A[a][x] = f_a(x) + fields of the same loop body; AA_POST: A[opp(a)][x] = f_post_a(x) of the
                                     // NEXT body's collision, fields of the current body, lid = rho plane of the previous macro()
    double *rho, *u, *v, *w, *up, *vp, *wp;
    double *lid;                     // rho(:,:,nz) for the moving-lid term, L3/bounce_back.f90:77-78
    double *scratch, *stage;
    long long stage_doubles, launches, bytes;
    cudaStream_t s;
    cudaEvent_t ev_t0, ev_t1;
};

static inline long long aa_ncell(const mglc_aa *h) { return (long long)h->g.nx * h->g.ny * h->g.nz; }
static int aa_use(mglc_aa *h) {
    if (!h) { set_error("mglc_aa: null handle"); return MGLC_E_INVALID; }
    MGLC_CUDA(cudaSetDevice(h->d.device));
    return MGLC_OK;
}
static int aa_malloc(mglc_aa *h, double **p, long long count) {
    if (cudaMalloc((void **)p, (size_t)count * sizeof(double)) != cudaSuccess) {
        (void)cudaGetLastError();
        *p = nullptr;
        set_error("mglc_aa: out of device memory (%.1f GB requested on top of %.1f GB)", count * 8 / 1e9, h->bytes / 1e9);
        return MGLC_E_NOMEM;
    }
    h->bytes += count * 8;
    return MGLC_OK;
}

extern "C" int mglc_aa_desc_init(mglc_aa_desc *d, int nx, int ny, int nz, double reynolds, double U0, double rho0) {
    if (!d || nx < 1 || ny < 1 || nz < 1 || !(reynolds > 0.0)) { set_error("mglc_aa_desc_init: bad arguments"); return MGLC_E_INVALID; }
    memset(d, 0, sizeof *d);
    d->n[0] = nx; d->n[1] = ny; d->n[2] = nz;
    d->arith = MGLC_ARITH_FAST; d->collision = MGLC_MRT_LID;
    d->tau = U0 * (double)nx / reynolds * 3.0 + 0.5;       // L3/commondata.f90:9
    d->U0 = U0; d->rho0 = rho0;
    return MGLC_OK;
}

extern "C" int mglc_aa_destroy(mglc_aa *h) {
    if (!h) return MGLC_OK;
    cudaSetDevice(h->d.device);
    if (h->s) cudaStreamSynchronize(h->s);
    double *bufs[] = {h->A, h->rho, h->u, h->v, h->w, h->up, h->vp, h->wp, h->lid, h->scratch, h->stage};
    for (double *p : bufs) cudaFree(p);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->s) cudaStreamDestroy(h->s);
    (void)cudaGetLastError();
    delete h;
    return MGLC_OK;
}

extern "C" int mglc_aa_create(mglc_aa **out, const mglc_aa_desc *d) {
    if (!out || !d) { set_error("mglc_aa_create: null argument"); return MGLC_E_INVALID; }
    if (d->n[0] < 1 || d->n[1] < 1 || d->n[2] < 1 || !(d->tau > 0.5) || (d->arith != MGLC_ARITH_FAST && d->arith != MGLC_ARITH_STRICT) ||
        (d->collision != MGLC_MRT_LID && d->collision != MGLC_BGK)) {
        set_error("mglc_aa_create: bad descriptor (%d x %d x %d, tau %g, arith %d, collision %d)", d->n[0], d->n[1], d->n[2], d->tau, d->arith, d->collision);
        return MGLC_E_INVALID;
    }
    MGLC_TRY(require_gpu());
    mglc_aa *h = new mglc_aa();
    memset(h, 0, sizeof *h);
    h->d = *d;
    h->g = make_geom(d->n[0], d->n[1], d->n[2]);
    for (int f = 0; f < 6; ++f) h->g.wall[f] = 1;          // one subdomain: every face is a wall of the global box
    h->g.lid = 1;
    h->p.Snu = 1.0 / d->tau;                                // L3/commondata.f90:42
    h->p.Sq = 8.0 * (2.0 * d->tau - 1.0) / (8.0 * d->tau - 1.0);
    h->p.U0 = d->U0; h->p.rho0 = d->rho0; h->p.bgk = d->collision == MGLC_BGK;
    auto fail = [&](int rc) { mglc_aa_destroy(h); return rc; };
    if (cudaSetDevice(d->device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", d->device); return fail(MGLC_E_CUDA); }
    if (cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&h->ev_t0) != cudaSuccess ||
        cudaEventCreate(&h->ev_t1) != cudaSuccess) { set_error("mglc_aa_create: stream/event creation failed"); return fail(MGLC_E_CUDA); }
    const long long n = aa_ncell(h);
    int rc;
    if ((rc = aa_malloc(h, &h->A, (long long)Q * h->g.sq)) || (rc = aa_malloc(h, &h->rho, n)) || (rc = aa_malloc(h, &h->u, n)) ||
        (rc = aa_malloc(h, &h->v, n)) || (rc = aa_malloc(h, &h->w, n)) || (rc = aa_malloc(h, &h->lid, (long long)h->g.nx * h->g.ny)) ||
        (rc = aa_malloc(h, &h->scratch, check_scratch_doubles()))) return fail(rc);
    // the halo ring is never read or written on this path; zero it once so that a download of the raw lattice is defined
    if (cudaMemsetAsync(h->A, 0, (size_t)Q * h->g.sq * sizeof(double), h->s) != cudaSuccess || cudaStreamSynchronize(h->s) != cudaSuccess)
        return fail(MGLC_E_CUDA);
    h->layout = AA_NATURAL;
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_aa_device_bytes(mglc_aa *h, long long *bytes) {
    if (!h || !bytes) return MGLC_E_INVALID;
    *bytes = h->bytes;
    return MGLC_OK;
}
extern "C" int mglc_aa_launch_count(mglc_aa *h, long long *n) {
    if (!h || !n) return MGLC_E_INVALID;
    *n = h->launches;
    return MGLC_OK;
}
extern "C" int mglc_aa_sync(mglc_aa *h) {
    MGLC_TRY(aa_use(h));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

static dim3 aa_grid_h(const mglc_aa *h) { return dim3((h->g.nx + 127) / 128, h->g.ny, h->g.nz); }
static int aa_ensure_stage(mglc_aa *h) {
    if (h->stage) return MGLC_OK;
    h->stage_doubles = 8LL << 20;   // 64 MiB
    return aa_malloc(h, &h->stage, h->stage_doubles);
}

// initial(): L3/initial.f90:55-73 (rho = rho0, u = U0 on the lid plane, f = feq)
extern "C" int mglc_aa_initial(mglc_aa *h) {
    MGLC_TRY(aa_use(h));
    h->launches += launch_initial(h->g, h->p, h->A, h->rho, h->u, h->v, h->w, h->s);
    if (h->up) {
        const size_t b = (size_t)aa_ncell(h) * sizeof(double);
        MGLC_CUDA(cudaMemsetAsync(h->up, 0, b, h->s)); MGLC_CUDA(cudaMemsetAsync(h->vp, 0, b, h->s)); MGLC_CUDA(cudaMemsetAsync(h->wp, 0, b, h->s));
    }
    h->layout = AA_NATURAL;
    return MGLC_OK;
}

// f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz) in the reference's layout; NULL = keep.  Leaves the NATURAL layout.
extern "C" int mglc_aa_upload(mglc_aa *h, const double *f, const double *rho, const double *u, const double *v, const double *w) {
    MGLC_TRY(aa_use(h));
    if (h->layout != AA_NATURAL && !f) { set_error("mglc_aa_upload: the lattice is between two streaming steps; upload f as well"); return MGLC_E_STATE; }
    if (f) {
        MGLC_TRY(aa_ensure_stage(h));
        const long long total = aa_ncell(h), chunk = h->stage_doubles / Q;
        for (long long c0 = 0; c0 < total; c0 += chunk) {
            const long long nc = std::min(chunk, total - c0);
            MGLC_CUDA(cudaMemcpyAsync(h->stage, f + c0 * Q, (size_t)nc * Q * sizeof(double), cudaMemcpyHostToDevice, h->s));
            h->launches += launch_aos_to_soa(h->g, Q, h->stage, h->A, c0, nc, 0, h->s);
        }
        h->layout = AA_NATURAL;
    }
    const size_t b = (size_t)aa_ncell(h) * sizeof(double);
    const double *src[4] = {rho, u, v, w};
    double *dst[4] = {h->rho, h->u, h->v, h->w};
    for (int q = 0; q < 4; ++q)
        if (src[q]) MGLC_CUDA(cudaMemcpyAsync(dst[q], src[q], b, cudaMemcpyHostToDevice, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return MGLC_OK;
}

extern "C" int mglc_aa_download_macro(mglc_aa *h, double *rho, double *u, double *v, double *w) {
    MGLC_TRY(aa_use(h));
    const size_t b = (size_t)aa_ncell(h) * sizeof(double);
    double *dst[4] = {rho, u, v, w};
    double *src[4] = {h->rho, h->u, h->v, h->w};
    for (int q = 0; q < 4; ++q)
        if (dst[q]) MGLC_CUDA(cudaMemcpyAsync(dst[q], src[q], b, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

// f as the reference holds it after the last loop body (pre-collision): a transposed copy in the NATURAL layout, a gather
// through streaming() + bounceback() in the POST layout
extern "C" int mglc_aa_download_f(mglc_aa *h, double *f) {
    MGLC_TRY(aa_use(h));
    if (!f) return MGLC_E_INVALID;
    MGLC_TRY(aa_ensure_stage(h));
    const long long total = aa_ncell(h), chunk = h->stage_doubles / Q;
    for (long long c0 = 0; c0 < total; c0 += chunk) {
        const long long nc = std::min(chunk, total - c0);
        if (h->layout == AA_NATURAL) h->launches += launch_soa_to_aos(h->g, Q, h->A, h->stage, c0, nc, 0, h->s);
        else {
            k_aa_gather_f<<<(unsigned)((nc + 127) / 128), 128, 0, h->s>>>(h->g, h->p, h->A, h->lid, c0, nc, h->stage);
            h->launches += 1;
        }
        MGLC_CUDA(cudaMemcpyAsync(f + c0 * Q, h->stage, (size_t)nc * Q * sizeof(double), cudaMemcpyDeviceToHost, h->s));
    }
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

// nsteps loop bodies (L3/main.f90:89-97): the schedule is aa_run() (lbm_aa.cuh).  rho,u,v,w afterwards are the reference's
// after the same number of loop bodies.
static int aa_step_impl(mglc_aa *h, int nsteps) {
    if (nsteps < 0) { set_error("mglc_aa_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    const bool st = h->d.arith == MGLC_ARITH_STRICT;
    h->launches += aa_run(h->layout, nsteps, [&](AaOp op) -> long long {
        switch (op) {
        case AA_OP_LID_PLANE:
            k_aa_lid_plane<<<(unsigned)(((long long)h->g.nx * h->g.ny + 255) / 256), 256, 0, h->s>>>(h->g, h->rho, h->lid);
            return 1;
        case AA_OP_COLLIDE0: return (st ? strict::launch_aa_collide0 : fast::launch_aa_collide0)(h->g, h->p, h->A, h->rho, h->u, h->v, h->w, h->s);
        case AA_OP_ODD: return (st ? strict::launch_aa_odd : fast::launch_aa_odd)(h->g, h->p, h->A, h->lid, h->s);
        case AA_OP_EVEN: return (st ? strict::launch_aa_even : fast::launch_aa_even)(h->g, h->p, h->A, h->lid, h->s);
        case AA_OP_MACRO_POST:
            k_aa_macro_post<<<aa_grid_h(h), 128, 0, h->s>>>(h->g, h->p, h->A, h->lid, h->rho, h->u, h->v, h->w);
            return 1;
        case AA_OP_MACRO: return launch_macro(h->g, h->A, h->rho, h->u, h->v, h->w, h->s);
        }
        return 0;
    });
    return MGLC_OK;
}
extern "C" int mglc_aa_step(mglc_aa *h, int nsteps) {
    MGLC_TRY(aa_use(h));
    MGLC_TRY(aa_step_impl(h, nsteps));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}
extern "C" int mglc_aa_step_timed(mglc_aa *h, int nsteps, float *ms) {
    MGLC_TRY(aa_use(h));
    if (!ms) return MGLC_E_INVALID;
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaEventRecord(h->ev_t0, h->s));
    MGLC_TRY(aa_step_impl(h, nsteps));
    MGLC_CUDA(cudaEventRecord(h->ev_t1, h->s));
    MGLC_CUDA(cudaEventSynchronize(h->ev_t1));
    MGLC_CUDA(cudaGetLastError());
    MGLC_CUDA(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
    return MGLC_OK;
}

// check(): L3/check.f90:12-34 (up, vp, wp are allocated on first use: 24 B/cell that a run without residual checks keeps free)
extern "C" int mglc_aa_check(mglc_aa *h, double *errorU) {
    MGLC_TRY(aa_use(h));
    if (!errorU) return MGLC_E_INVALID;
    if (!h->up) {
        const long long n = aa_ncell(h);
        MGLC_TRY(aa_malloc(h, &h->up, n)); MGLC_TRY(aa_malloc(h, &h->vp, n)); MGLC_TRY(aa_malloc(h, &h->wp, n));
        MGLC_CUDA(cudaMemsetAsync(h->up, 0, (size_t)n * 8, h->s)); MGLC_CUDA(cudaMemsetAsync(h->vp, 0, (size_t)n * 8, h->s));
        MGLC_CUDA(cudaMemsetAsync(h->wp, 0, (size_t)n * 8, h->s));          // up = vp = wp = 0, L3/initial.f90:50-52
    }
    h->launches += launch_check(h->g, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->s);
    double e[2];
    MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaGetLastError());
    *errorU = sqrt(e[0]) / sqrt(e[1]);
    return MGLC_OK;
}
