// lbm_aa.cu -- the D3Q19 lid-driven cavity loop (L3/main.f90:85-103) on ONE lattice (AA-pattern storage)
// behind the mglc_aa_* entry points of mglc.h: the same arithmetic and the same results as mglc_lbm_step (bit-identical in
// strict mode), but 19 x 8 B per cell of lattice instead of 2 x 19 x 8 B, for the largest lattices one GPU can hold.
// One subdomain (mglc_aa_create: all six faces are walls), or the blocks of a decomposed lattice inside one process
// (mglc_aa_group_*): every launch also stores what its neighbours need straight into their lattices (lbm_aa_kernels.inl), and
// neighbouring launches are ordered by events, one epoch per launch.
// This file holds the strict build of the kernels (-fmad=false), the order-preserving kernels and the host side.
#define MGLC_NS strict
#define MGLC_STRICT 1
#include "lbm_aa_kernels.inl"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

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
    // a block of a decomposed lattice (mglc_aa_group_*): position in the process grid, the neighbours in the 18 directions of
    // ex_sendrecv.f90's messages (index = face 0..5 or edge population 7..18, -1 = none), their lattices as the kernels see them
    struct mglc_aa_group *group;
    int rank, coords[3], dims[3], start[3], nbr[19];
    PeerTable peers;                 // peers.mask != 0: passed to the PEER build of the kernels
    cudaEvent_t ev_done[2];          // group: this block's launch of epoch e has finished: ev_done[e & 1]
    // one process per block (mglc_aa_create_comm): the neighbours' lattices are CUDA IPC mappings, launches are ordered by the
    // flag words of k_halo_signal / k_halo_wait (word = 2 x the number of launches the block has finished)
    mglc_comm *comm;
    unsigned long long *flags;
    int *d_err;
    SyncTable sync;
    long long epoch;
    std::vector<void *> *ipc_opened;
};
struct mglc_aa_group {
    std::vector<mglc_aa *> m;
    mglc_aa_desc global;
    int dims[3];
    int layout;                      // one layout for all blocks: they advance launch by launch together
    long long epoch;                 // launches issued per block so far
};

static inline long long aa_ncell(const mglc_aa *h) { return (long long)h->g.nx * h->g.ny * h->g.nz; }
static int aa_use(mglc_aa *h) {
    if (!h) { set_error("mglc_aa: null handle"); return MGLC_E_INVALID; }
    MGLC_CUDA(cudaSetDevice(h->d.device));
    return MGLC_OK;
}
// the barrier's sticky error word (k_halo_wait): a run whose neighbours fell out of step never comes back as MGLC_OK
static int aa_comm_status(mglc_aa *h) {
    if (!h->comm || !h->d_err) return MGLC_OK;
    int e = 0;
    MGLC_CUDA(cudaMemcpy(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e) { set_error("mglc_aa: a neighbour did not reach the barrier within %.0f s (MGLC_HALO_TIMEOUT_S; ranks out of step?)", halo_timeout_seconds()); return MGLC_E_STATE; }
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
    if (h->comm && h->sync.mask && h->s && h->epoch > 0) {
        // the neighbours may still be storing into this block: wait for their last launch (bounded by the barrier's timeout)
        int e = 0;
        if (cudaMemcpy(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost) == cudaSuccess && !e) launch_halo_wait(h->sync, (unsigned long long)h->epoch * 2, h->d_err, h->s);
    }
    if (h->s) cudaStreamSynchronize(h->s);
    if (h->ipc_opened) { for (void *p : *h->ipc_opened) cudaIpcCloseMemHandle(p); delete h->ipc_opened; }
    cudaFree(h->flags); cudaFree(h->d_err);
    double *bufs[] = {h->A, h->rho, h->u, h->v, h->w, h->up, h->vp, h->wp, h->lid, h->scratch, h->stage};
    for (double *p : bufs) cudaFree(p);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    for (cudaEvent_t e : h->ev_done) if (e) cudaEventDestroy(e);
    if (h->s) cudaStreamDestroy(h->s);
    (void)cudaGetLastError();
    delete h;
    return MGLC_OK;
}

// wall[f]: face f (+x,-x,+y,-y,+z,-z) of this block is a wall of the global box; lid: the block touches the moving lid
static int aa_create_impl(mglc_aa **out, const mglc_aa_desc *d, const int *wall, int lid) {
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
    for (int f = 0; f < 6; ++f) h->g.wall[f] = wall[f];
    h->g.lid = lid;
    for (int &r : h->nbr) r = -1;
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
    // one subdomain: the halo ring is never read or written; a decomposed block: only the slots the neighbours store into are
    // read.  Zero it once so that a download of the raw lattice is defined
    if (cudaMemsetAsync(h->A, 0, (size_t)Q * h->g.sq * sizeof(double), h->s) != cudaSuccess || cudaStreamSynchronize(h->s) != cudaSuccess)
        return fail(MGLC_E_CUDA);
    h->layout = AA_NATURAL;
    *out = h;
    return MGLC_OK;
}
extern "C" int mglc_aa_create(mglc_aa **out, const mglc_aa_desc *d) {
    const int wall[6] = {1, 1, 1, 1, 1, 1};                 // one subdomain: every face is a wall of the global box
    return aa_create_impl(out, d, wall, 1);
}
// entry points that move one block on its own are refused on a member of a group: the blocks advance together
static int aa_single(mglc_aa *h, const char *what) {
    if (h && h->group) { set_error("%s: this handle is a block of a group; use the mglc_aa_group_* call", what); return MGLC_E_STATE; }
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
    return aa_comm_status(h);
}

// make this block's stream wait for the last launch of each neighbour (what they stored into this block is complete, and
// they have finished reading what the next launch of this block will overwrite in them)
static int aa_wait_neighbours(mglc_aa *h) {
    if (h->comm) {
        if (h->sync.mask && h->epoch > 0) h->launches += launch_halo_wait(h->sync, (unsigned long long)h->epoch * 2, h->d_err, h->s);
        return MGLC_OK;
    }
    mglc_aa_group *G = h->group;
    if (!G || G->epoch == 0) return MGLC_OK;
    for (int d = 0; d < 19; ++d) {
        if (h->nbr[d] < 0) continue;
        bool seen = false;
        for (int e = 0; e < d; ++e) seen = seen || h->nbr[e] == h->nbr[d];
        if (!seen) MGLC_CUDA(cudaStreamWaitEvent(h->s, G->m[h->nbr[d]]->ev_done[(G->epoch - 1) & 1], 0));
    }
    return MGLC_OK;
}

static dim3 aa_grid_h(const mglc_aa *h) { return dim3((h->g.nx + 127) / 128, h->g.ny, h->g.nz); }
static int aa_ensure_stage(mglc_aa *h) {
    if (h->stage) return MGLC_OK;
    h->stage_doubles = 8LL << 20;   // 64 MiB
    return aa_malloc(h, &h->stage, h->stage_doubles);
}

// initial(): L3/initial.f90:55-73 (rho = rho0, u = U0 on the lid plane, f = feq)
static int aa_initial_impl(mglc_aa *h) {
    MGLC_TRY(aa_use(h));
    if (h->comm) MGLC_TRY(aa_wait_neighbours(h));      // their last launch may still be storing into this block
    h->launches += launch_initial(h->g, h->p, h->A, h->rho, h->u, h->v, h->w, h->s);
    if (h->up) {
        const size_t b = (size_t)aa_ncell(h) * sizeof(double);
        MGLC_CUDA(cudaMemsetAsync(h->up, 0, b, h->s)); MGLC_CUDA(cudaMemsetAsync(h->vp, 0, b, h->s)); MGLC_CUDA(cudaMemsetAsync(h->wp, 0, b, h->s));
    }
    h->layout = AA_NATURAL;
    return MGLC_OK;
}
extern "C" int mglc_aa_initial(mglc_aa *h) {
    MGLC_TRY(aa_single(h, "mglc_aa_initial"));
    return aa_initial_impl(h);
}

// f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz) in the reference's layout; NULL = keep.  Leaves the NATURAL layout.
// (a block of a group: only while the group is in the NATURAL layout -- after mglc_aa_group_initial or an even number of
// loop bodies -- since all blocks share one layout)
extern "C" int mglc_aa_upload(mglc_aa *h, const double *f, const double *rho, const double *u, const double *v, const double *w) {
    MGLC_TRY(aa_use(h));
    if (h->group && h->layout != AA_NATURAL) { set_error("mglc_aa_upload: the group is between two streaming steps"); return MGLC_E_STATE; }
    if (h->comm) MGLC_TRY(aa_wait_neighbours(h));
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
    return aa_comm_status(h);
}

// f as the reference holds it after the last loop body (pre-collision): a transposed copy in the NATURAL layout, a gather
// through streaming() + bounceback() in the POST layout
extern "C" int mglc_aa_download_f(mglc_aa *h, double *f) {
    MGLC_TRY(aa_use(h));
    if (!f) return MGLC_E_INVALID;
    MGLC_TRY(aa_ensure_stage(h));
    MGLC_TRY(aa_wait_neighbours(h));         // POST layout of a block: the gather pulls what the neighbours parked in its halos
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
    return aa_comm_status(h);
}

// one launch of the schedule aa_run() (lbm_aa.cuh) on one block
static long long aa_launch_op(mglc_aa *h, AaOp op) {
    const bool st = h->d.arith == MGLC_ARITH_STRICT;
    const PeerTable *pt = h->peers.mask ? &h->peers : nullptr;
    switch (op) {
    case AA_OP_LID_PLANE:
        k_aa_lid_plane<<<(unsigned)(((long long)h->g.nx * h->g.ny + 255) / 256), 256, 0, h->s>>>(h->g, h->rho, h->lid);
        return 1;
    case AA_OP_COLLIDE0: return (st ? strict::launch_aa_collide0 : fast::launch_aa_collide0)(h->g, h->p, h->A, h->rho, h->u, h->v, h->w, h->s, pt);
    case AA_OP_ODD: return (st ? strict::launch_aa_odd : fast::launch_aa_odd)(h->g, h->p, h->A, h->lid, h->s, pt);
    case AA_OP_EVEN: return (st ? strict::launch_aa_even : fast::launch_aa_even)(h->g, h->p, h->A, h->lid, h->s, pt);
    case AA_OP_MACRO_POST:
        k_aa_macro_post<<<aa_grid_h(h), 128, 0, h->s>>>(h->g, h->p, h->A, h->lid, h->rho, h->u, h->v, h->w);
        return 1;
    case AA_OP_MACRO: return launch_macro(h->g, h->A, h->rho, h->u, h->v, h->w, h->s);
    }
    return 0;
}
// nsteps loop bodies (L3/main.f90:89-97): the schedule is aa_run() (lbm_aa.cuh).  rho,u,v,w afterwards are the reference's
// after the same number of loop bodies.
static int aa_step_impl(mglc_aa *h, int nsteps) {
    if (nsteps < 0) { set_error("mglc_aa_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    if (!h->comm || !h->sync.mask) {
        h->launches += aa_run(h->layout, nsteps, [&](AaOp op) -> long long { return aa_launch_op(h, op); });
        return MGLC_OK;
    }
    // one process per block: every launch is one epoch of the neighbour barrier (all ranks run the same schedule)
    h->launches += aa_run(h->layout, nsteps, [&](AaOp op) -> long long {
        long long n = 0;
        if (h->epoch > 0) n += launch_halo_wait(h->sync, (unsigned long long)h->epoch * 2, h->d_err, h->s);
        n += aa_launch_op(h, op);
        ++h->epoch;
        return n + launch_halo_signal(h->sync, (unsigned long long)h->epoch * 2, h->s);
    });
    return MGLC_OK;
}
extern "C" int mglc_aa_step(mglc_aa *h, int nsteps) {
    MGLC_TRY(aa_single(h, "mglc_aa_step"));
    MGLC_TRY(aa_use(h));
    MGLC_TRY(aa_step_impl(h, nsteps));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}
extern "C" int mglc_aa_step_timed(mglc_aa *h, int nsteps, float *ms) {
    MGLC_TRY(aa_single(h, "mglc_aa_step_timed"));
    MGLC_TRY(aa_use(h));
    if (!ms) return MGLC_E_INVALID;
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaEventRecord(h->ev_t0, h->s));
    MGLC_TRY(aa_step_impl(h, nsteps));
    MGLC_CUDA(cudaEventRecord(h->ev_t1, h->s));
    MGLC_CUDA(cudaEventSynchronize(h->ev_t1));
    MGLC_CUDA(cudaGetLastError());
    MGLC_CUDA(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
    return aa_comm_status(h);
}

// check(): L3/check.f90:12-34 (up, vp, wp are allocated on first use: 24 B/cell that a run without residual checks keeps free)
static int aa_check_partial(mglc_aa *h, double (&e)[2]) {
    MGLC_TRY(aa_use(h));
    if (!h->up) {
        const long long n = aa_ncell(h);
        MGLC_TRY(aa_malloc(h, &h->up, n)); MGLC_TRY(aa_malloc(h, &h->vp, n)); MGLC_TRY(aa_malloc(h, &h->wp, n));
        MGLC_CUDA(cudaMemsetAsync(h->up, 0, (size_t)n * 8, h->s)); MGLC_CUDA(cudaMemsetAsync(h->vp, 0, (size_t)n * 8, h->s));
        MGLC_CUDA(cudaMemsetAsync(h->wp, 0, (size_t)n * 8, h->s));          // up = vp = wp = 0, L3/initial.f90:50-52
    }
    h->launches += launch_check(h->g, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->s);
    if (h->comm && h->comm->nranks > 1) MGLC_NCCL(ncclAllReduce(h->scratch, h->scratch, 2, ncclDouble, ncclSum, h->comm->nccl, h->s));   // check.f90:27-28
    MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaGetLastError());
    return aa_comm_status(h);
}
extern "C" int mglc_aa_check(mglc_aa *h, double *errorU) {
    if (!errorU) return MGLC_E_INVALID;
    MGLC_TRY(aa_single(h, "mglc_aa_check"));
    double e[2];
    MGLC_TRY(aa_check_partial(h, e));
    *errorU = sqrt(e[0]) / sqrt(e[1]);
    return MGLC_OK;
}

// ================= a decomposed lattice inside one process (SURVEY 8f row 4 + 8e) =================
// The blocks of mpi_starts/decompose_1d (L3/main.f90:33-63) on one or several devices.  No halo message is ever packed: a
// launch that parks post-collision populations also stores those of its boundary cells into the neighbours' halo cells, and
// the odd launch pushes across a face straight into the neighbour's own cells (lbm_aa_kernels.inl).  Ordering: launch e+1 of a
// block waits for launch e of its (up to 18) neighbours -- the same neighbour barrier as the direct-halo path of
// mglc_group_step, here through events.
extern "C" int mglc_aa_group_destroy(mglc_aa_group *G) {
    if (!G) return MGLC_OK;
    for (mglc_aa *h : G->m) { if (h) { cudaSetDevice(h->d.device); cudaStreamSynchronize(h->s); } }     // nobody still writes into a block that goes away
    for (mglc_aa *h : G->m) mglc_aa_destroy(h);
    delete G;
    return MGLC_OK;
}
extern "C" int mglc_aa_group_create(mglc_aa_group **out, const mglc_aa_desc *gd, int nranks, const int *dims_or_null, const int *devices_or_null) {
    if (!out || !gd || nranks < 1) { set_error("mglc_aa_group_create: bad arguments"); return MGLC_E_INVALID; }
    MGLC_TRY(require_gpu());
    mglc_aa_group *G = new mglc_aa_group();
    G->global = *gd;
    G->layout = AA_NATURAL;
    G->epoch = 0;
    if (dims_or_null && dims_or_null[0] > 0) for (int q = 0; q < 3; ++q) G->dims[q] = dims_or_null[q];
    else { G->dims[0] = G->dims[1] = G->dims[2] = 0; mglc_dims_create(nranks, G->dims); }
    auto fail = [&](int rc) { mglc_aa_group_destroy(G); return rc; };
    if (G->dims[0] < 1 || G->dims[1] < 1 || G->dims[2] < 1 || G->dims[0] * G->dims[1] * G->dims[2] != nranks) {
        set_error("mglc_aa_group_create: dims %dx%dx%d do not multiply to %d", G->dims[0], G->dims[1], G->dims[2], nranks);
        return fail(MGLC_E_INVALID);
    }
    for (int r = 0; r < nranks; ++r) {
        mglc_aa_desc d = *gd;
        int coords[3], start[3], wall[6];
        mglc_cart_coords(G->dims, r, coords);
        for (int q = 0; q < 3; ++q) {
            const int rc = mglc_decompose_1d(gd->n[q], coords[q], G->dims[q], &d.n[q], &start[q]);
            if (rc) return fail(rc);
            wall[2 * q] = coords[q] == G->dims[q] - 1;
            wall[2 * q + 1] = coords[q] == 0;
        }
        d.device = devices_or_null ? devices_or_null[r] : gd->device;
        mglc_aa *h = nullptr;
        const int rc = aa_create_impl(&h, &d, wall, wall[4]);
        if (rc) return fail(rc);
        G->m.push_back(h);
        h->group = G; h->rank = r;
        for (int q = 0; q < 3; ++q) { h->coords[q] = coords[q]; h->dims[q] = G->dims[q]; h->start[q] = start[q]; }
        for (int dd = 0; dd < 19; ++dd) {
            if (dd == 6) continue;
            const int e[3] = {dd < 6 ? (dd >> 1 == 0 ? 1 - 2 * (dd & 1) : 0) : h_ex[dd], dd < 6 ? (dd >> 1 == 1 ? 1 - 2 * (dd & 1) : 0) : h_ey[dd],
                              dd < 6 ? (dd >> 1 == 2 ? 1 - 2 * (dd & 1) : 0) : h_ez[dd]};
            const int c[3] = {coords[0] + e[0], coords[1] + e[1], coords[2] + e[2]};
            mglc_cart_rank(G->dims, c, &h->nbr[dd]);
        }
        for (cudaEvent_t &e : h->ev_done) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(MGLC_E_CUDA);
    }
    // the neighbours' lattices must be addressable from every block's device
    for (mglc_aa *a : G->m)
        for (mglc_aa *b : G->m)
            if (a->d.device != b->d.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->d.device, b->d.device);
                if (!can) { set_error("mglc_aa_group_create: device %d cannot address device %d (no peer access)", a->d.device, b->d.device); return fail(MGLC_E_CUDA); }
                cudaSetDevice(a->d.device); cudaDeviceEnablePeerAccess(b->d.device, 0); (void)cudaGetLastError();
            }
    for (mglc_aa *h : G->m) {
        PeerTable t;
        memset(&t, 0, sizeof t);
        for (int dd = 0; dd < 19; ++dd) {
            if (dd == 6 || h->nbr[dd] < 0) continue;
            const mglc_aa *n = G->m[h->nbr[dd]];
            t.mask |= 1u << dd;
            t.F[dd] = n->A; t.sy[dd] = n->g.sy; t.sz[dd] = n->g.sz; t.sq[dd] = n->g.sq;
            t.n[dd][0] = n->g.nx; t.n[dd][1] = n->g.ny; t.n[dd][2] = n->g.nz;
        }
        h->peers = t;                        // mask 0 (a 1-block group): the one-subdomain kernels
    }
    *out = G;
    return MGLC_OK;
}
extern "C" int mglc_aa_group_size(mglc_aa_group *G, int *n) { if (!G || !n) return MGLC_E_INVALID; *n = (int)G->m.size(); return MGLC_OK; }
extern "C" int mglc_aa_group_dims(mglc_aa_group *G, int *dims) { if (!G || !dims) return MGLC_E_INVALID; for (int q = 0; q < 3; ++q) dims[q] = G->dims[q]; return MGLC_OK; }
extern "C" int mglc_aa_group_rank(mglc_aa_group *G, int r, mglc_aa **h) {
    if (!G || !h || r < 0 || r >= (int)G->m.size()) return MGLC_E_INVALID;
    *h = G->m[r];
    return MGLC_OK;
}
// local size and 0-based global offset of a block (decompose_1d, L3/main.f90:247-263); a plain handle: the whole lattice
extern "C" int mglc_aa_get_block(mglc_aa *h, int *ln, int *start) {
    if (!h || !ln || !start) return MGLC_E_INVALID;
    ln[0] = h->g.nx; ln[1] = h->g.ny; ln[2] = h->g.nz;
    for (int q = 0; q < 3; ++q) start[q] = (h->group || h->comm) ? h->start[q] : 0;
    return MGLC_OK;
}
extern "C" int mglc_aa_group_initial(mglc_aa_group *G) {
    if (!G) return MGLC_E_INVALID;
    // a block's halos and boundary cells may still be the target of a neighbour's launch: order initial() after all of them
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaStreamSynchronize(h->s)); }
    for (mglc_aa *h : G->m) MGLC_TRY(aa_initial_impl(h));
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaStreamSynchronize(h->s)); }
    G->layout = AA_NATURAL;
    return MGLC_OK;
}
static int aa_group_step_impl(mglc_aa_group *G, int nsteps) {
    if (nsteps < 0) { set_error("mglc_aa_group_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    for (mglc_aa *h : G->m)
        if (h->layout != G->layout) { set_error("mglc_aa_group_step: block %d was moved on its own", h->rank); return MGLC_E_STATE; }
    int rc = MGLC_OK;
    aa_run(G->layout, nsteps, [&](AaOp op) -> long long {
        for (mglc_aa *h : G->m) {
            if (rc) break;
            if ((rc = aa_use(h)) || (rc = aa_wait_neighbours(h))) break;
            h->launches += aa_launch_op(h, op);
            if (cudaEventRecord(h->ev_done[G->epoch & 1], h->s) != cudaSuccess) { set_error("mglc_aa_group_step: cudaEventRecord failed"); rc = MGLC_E_CUDA; }
        }
        ++G->epoch;
        return 0;
    });
    for (mglc_aa *h : G->m) h->layout = G->layout;
    return rc;
}
extern "C" int mglc_aa_group_step(mglc_aa_group *G, int nsteps) {
    if (!G) return MGLC_E_INVALID;
    MGLC_TRY(aa_group_step_impl(G, nsteps));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}
// device time of nsteps loop bodies: the longest of the blocks' streams (every block starts from an idle group)
extern "C" int mglc_aa_group_step_timed(mglc_aa_group *G, int nsteps, float *ms) {
    if (!G || !ms) return MGLC_E_INVALID;
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaStreamSynchronize(h->s)); }
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaEventRecord(h->ev_t0, h->s)); }
    MGLC_TRY(aa_group_step_impl(G, nsteps));
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaEventRecord(h->ev_t1, h->s)); }
    *ms = 0.0f;
    for (mglc_aa *h : G->m) {
        MGLC_TRY(aa_use(h));
        MGLC_CUDA(cudaEventSynchronize(h->ev_t1));
        float t = 0.0f;
        MGLC_CUDA(cudaEventElapsedTime(&t, h->ev_t0, h->ev_t1));
        *ms = std::max(*ms, t);
    }
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}
// check(): the two sums of L3/check.f90:12-34 per block, MPI_Allreduce(SUM) as a rank-ordered host sum
extern "C" int mglc_aa_group_check(mglc_aa_group *G, double *errorU) {
    if (!G || !errorU) return MGLC_E_INVALID;
    double t[2] = {0.0, 0.0};
    for (mglc_aa *h : G->m) {
        double e[2];
        MGLC_TRY(aa_check_partial(h, e));
        t[0] += e[0]; t[1] += e[1];
    }
    *errorU = sqrt(t[0]) / sqrt(t[1]);
    return MGLC_OK;
}
extern "C" int mglc_aa_group_sync(mglc_aa_group *G) {
    if (!G) return MGLC_E_INVALID;
    for (mglc_aa *h : G->m) { MGLC_TRY(aa_use(h)); MGLC_CUDA(cudaStreamSynchronize(h->s)); }
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

// ================= a decomposed lattice, one process per block (torchrun / mpirun) =================
// The same kernels and the same launch-by-launch schedule as the group above; the neighbours' lattices are CUDA IPC mappings
// (exchanged once through the communicator), the launches are ordered by the flag words of k_halo_signal / k_halo_wait.
// There is no message transport to fall back to on this path: without IPC peer mappings creation fails.
namespace {
struct AaIpcRecord { cudaIpcMemHandle_t A, flags; int ln[3]; };
}
extern "C" int mglc_aa_create_comm(mglc_aa **out, const mglc_aa_desc *gd, mglc_comm *comm, const int *dims_or_null) {
    if (!out || !gd || !comm) { set_error("mglc_aa_create_comm: null argument"); return MGLC_E_INVALID; }
    const int P = comm->nranks, r = comm->rank;
    int dims[3] = {0, 0, 0};
    if (dims_or_null && dims_or_null[0] > 0) for (int q = 0; q < 3; ++q) dims[q] = dims_or_null[q];
    else mglc_dims_create(P, dims);
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || dims[0] * dims[1] * dims[2] != P) {
        set_error("mglc_aa_create_comm: dims %dx%dx%d do not multiply to %d ranks", dims[0], dims[1], dims[2], P);
        return MGLC_E_INVALID;
    }
    mglc_aa_desc d = *gd;
    int coords[3], start[3], wall[6];
    mglc_cart_coords(dims, r, coords);
    for (int q = 0; q < 3; ++q) {
        MGLC_TRY(mglc_decompose_1d(gd->n[q], coords[q], dims[q], &d.n[q], &start[q]));
        wall[2 * q] = coords[q] == dims[q] - 1;
        wall[2 * q + 1] = coords[q] == 0;
    }
    d.device = comm->device;
    mglc_aa *h = nullptr;
    MGLC_TRY(aa_create_impl(&h, &d, wall, wall[4]));
    auto fail = [&](int rc) { mglc_aa_destroy(h); return rc; };
    h->comm = comm; h->rank = r;
    for (int q = 0; q < 3; ++q) { h->coords[q] = coords[q]; h->dims[q] = dims[q]; h->start[q] = start[q]; }
    for (int dd = 0; dd < 19; ++dd) {
        if (dd == 6) continue;
        const int e[3] = {dd < 6 ? (dd >> 1 == 0 ? 1 - 2 * (dd & 1) : 0) : h_ex[dd], dd < 6 ? (dd >> 1 == 1 ? 1 - 2 * (dd & 1) : 0) : h_ey[dd],
                          dd < 6 ? (dd >> 1 == 2 ? 1 - 2 * (dd & 1) : 0) : h_ez[dd]};
        const int c[3] = {coords[0] + e[0], coords[1] + e[1], coords[2] + e[2]};
        mglc_cart_rank(dims, c, &h->nbr[dd]);
    }
    if (cudaMalloc((void **)&h->flags, 32 * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc((void **)&h->d_err, sizeof(int)) != cudaSuccess ||
        cudaMemset(h->flags, 0, 32 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(h->d_err, 0, sizeof(int)) != cudaSuccess) return fail(MGLC_E_NOMEM);
    if (P == 1) { *out = h; return MGLC_OK; }
    // exchange the IPC handles of the lattices and the barrier words (collective)
    AaIpcRecord mine;
    memset(&mine, 0, sizeof mine);
    int ok = 1;
    const unsigned long long magic = 0x6d67616100000000ull + (unsigned long long)r;
    ok &= cudaMemcpy(h->flags + 31, &magic, sizeof magic, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaIpcGetMemHandle(&mine.A, h->A) == cudaSuccess;
    ok &= cudaIpcGetMemHandle(&mine.flags, h->flags) == cudaSuccess;
    (void)cudaGetLastError();
    mine.ln[0] = h->g.nx; mine.ln[1] = h->g.ny; mine.ln[2] = h->g.nz;
    char *dev_all = nullptr;
    std::vector<AaIpcRecord> all(P);
    {
        int rc = MGLC_OK;
        auto step = [&]() -> int {
            MGLC_CUDA(cudaMalloc((void **)&dev_all, (size_t)(P + 1) * sizeof(AaIpcRecord)));
            MGLC_CUDA(cudaMemcpy(dev_all + (size_t)P * sizeof(AaIpcRecord), &mine, sizeof mine, cudaMemcpyHostToDevice));
            MGLC_NCCL(ncclAllGather(dev_all + (size_t)P * sizeof(AaIpcRecord), dev_all, sizeof(AaIpcRecord), ncclChar, comm->nccl, h->s));
            MGLC_CUDA(cudaMemcpyAsync(all.data(), dev_all, (size_t)P * sizeof(AaIpcRecord), cudaMemcpyDeviceToHost, h->s));
            MGLC_CUDA(cudaStreamSynchronize(h->s));
            return MGLC_OK;
        };
        rc = step();
        cudaFree(dev_all);
        if (rc) return fail(rc);
    }
    h->ipc_opened = new std::vector<void *>();
    struct View { double *A; unsigned long long *flags; };
    std::vector<View> by_rank(P, View{nullptr, nullptr});
    std::vector<char> have(P, 0);
    auto open = [&](const cudaIpcMemHandle_t &hd, void **p) {
        if (cudaIpcOpenMemHandle(p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); *p = nullptr; return 0; }
        h->ipc_opened->push_back(*p);
        return 1;
    };
    memset(&h->peers, 0, sizeof h->peers);
    memset(&h->sync, 0, sizeof h->sync);
    h->peers.err = h->d_err;
    for (int dd = 0; dd < 19 && ok; ++dd) {
        const int n = h->nbr[dd];
        if (dd == 6 || n < 0) continue;
        if (!have[n]) {
            ok &= open(all[n].A, (void **)&by_rank[n].A);
            if (ok) ok &= open(all[n].flags, (void **)&by_rank[n].flags);
            if (ok) {           // the mapping must start at the neighbour's own pointer, not at some enclosing block
                unsigned long long seen = 0;
                ok &= cudaMemcpy(&seen, by_rank[n].flags + 31, sizeof seen, cudaMemcpyDeviceToHost) == cudaSuccess &&
                      seen == 0x6d67616100000000ull + (unsigned long long)n;
                (void)cudaGetLastError();
            }
            have[n] = 1;
        }
        if (!ok) break;
        const Geom pg = make_geom(all[n].ln[0], all[n].ln[1], all[n].ln[2]);
        h->peers.mask |= 1u << dd;
        h->peers.F[dd] = by_rank[n].A; h->peers.sy[dd] = pg.sy; h->peers.sz[dd] = pg.sz; h->peers.sq[dd] = pg.sq;
        h->peers.n[dd][0] = pg.nx; h->peers.n[dd][1] = pg.ny; h->peers.n[dd][2] = pg.nz;
        h->sync.mask |= 1u << dd;
        h->sync.signal[dd] = by_rank[n].flags + (dd < 6 ? (dd ^ 1) : h_opp[dd]);     // the neighbour sees me in the opposite direction
        h->sync.wait[dd] = h->flags + dd;
    }
    // collective verdict: everybody or nobody
    {
        int *dev_ok = nullptr, rc = MGLC_OK;
        auto step = [&]() -> int {
            MGLC_CUDA(cudaMalloc((void **)&dev_ok, sizeof(int)));
            MGLC_CUDA(cudaMemcpy(dev_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
            MGLC_NCCL(ncclAllReduce(dev_ok, dev_ok, 1, ncclInt, ncclMin, comm->nccl, h->s));
            MGLC_CUDA(cudaMemcpyAsync(&ok, dev_ok, sizeof ok, cudaMemcpyDeviceToHost, h->s));
            MGLC_CUDA(cudaStreamSynchronize(h->s));
            return MGLC_OK;
        };
        rc = step();
        cudaFree(dev_ok);
        if (rc) return fail(rc);
    }
    if (!ok) {
        h->sync.mask = 0;      // nothing to wait for in destroy
        set_error("mglc_aa_create_comm: the neighbours' lattices cannot be mapped (CUDA IPC / peer access); the single-lattice path has no message transport -- use mglc_lbm_create");
        return fail(MGLC_E_STATE);
    }
    *out = h;
    return MGLC_OK;
}
