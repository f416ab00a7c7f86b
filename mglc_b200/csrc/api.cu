// api.cu -- the device half of the C ABI (include/mglc.h): subdomain handles, host<->device
// transfers in the reference's array layout, the per-subroutine entry points, the fused step loop,
// halo exchange over NCCL (one process per GPU) or device-to-device copies (P subdomains in one
// process), and timing / launch accounting.  There is no CPU path: without a device every entry
// point here fails with MGLC_E_NOGPU.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "halo.cuh"

using namespace mglc;

struct mglc_lbm {
    mglc_lbm_desc d;
    Geom g;
    LbmParams p;
    int rank, nranks;
    double *buf[2];          // two halo'd SoA lattices; buf[cur] = f (pre-collision), buf[cur^1] = f_post
    int cur;
    // rotated: a fused run is in flight -- buf[cur^1] already holds the NEXT step's post-collision
    // populations (before exchange), buf[cur] the last step's f_post with its halos, and only the lid
    // plane of rho is current.  canonicalise() turns this back into the reference's state (f, rho,u,v,w).
    int rotated;
    double *rho, *u, *v, *w, *up, *vp, *wp;
    // lid plane (k = nz) of rho for the moving-lid term, L3/bounce_back.f90:77-78.  A fused launch reads
    // the plane the previous macro() left (lid_next) and writes this step's into the other side buffer,
    // so the plane it read (lid_last_in) stays intact for canonicalise().  In the reference's state
    // lid_next is simply the top plane of the rho field.
    double *rho_lid[2];
    const double *lid_next, *lid_last_in;
    double *scratch;         // check() partial sums
    cudaStream_t s, s_comm;
    cudaEvent_t ev_packed, ev_copied, ev_t0, ev_t1, ev_shell, ev_halo;
    long long launches;
    long long bytes;
    Msg msgs[18];
    int nmsgs;
    bool has_neighbors;
    mglc_comm *comm;
    mglc_group *group;
    double *stage;           // device staging for AoS<->SoA transposes
    long long stage_doubles;
    // optional per-launch timing of the fused kernel (CUDA events on the launching stream)
    int profiling;
    std::vector<cudaEvent_t> *prof_ev;   // pairs (start, stop), reused round-robin
    int prof_used;
    double prof_ms;
    long long prof_launches;
};

struct mglc_group {
    std::vector<mglc_lbm *> r;
    mglc_lbm_desc global;
};

// ---------------------------------------------------------------------------------------------------
static int use(mglc_lbm *h) {
    if (!h) { set_error("null handle"); return MGLC_E_INVALID; }
    MGLC_CUDA(cudaSetDevice(h->d.device));
    return MGLC_OK;
}
template <class T> static int dmalloc(mglc_lbm *h, T **p, long long count) {
    MGLC_CUDA(cudaMalloc((void **)p, (size_t)count * sizeof(T)));
    h->bytes += count * (long long)sizeof(T);
    return MGLC_OK;
}
static inline long long ncell(const mglc_lbm *h) { return (long long)h->g.nx * h->g.ny * h->g.nz; }
static inline double *F_(mglc_lbm *h) { return h->buf[h->cur]; }
static inline double *Fpost_(mglc_lbm *h) { return h->buf[h->cur ^ 1]; }
static inline bool strict_(const mglc_lbm *h) { return h->d.arith == MGLC_ARITH_STRICT; }

extern "C" int mglc_device_count(int *n) {
    if (!n) return MGLC_E_INVALID;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); *n = 0; }
    return MGLC_OK;
}

// ---- communicator -----------------------------------------------------------------------------------
extern "C" int mglc_comm_unique_id(char id[128]) {
    if (!id) return MGLC_E_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    MGLC_NCCL(ncclGetUniqueId(&uid));
    memcpy(id, &uid, 128);
    return MGLC_OK;
}
extern "C" int mglc_comm_init_rank(mglc_comm **c, const char id[128], int nranks, int rank, int device) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("mglc_comm_init_rank: bad arguments"); return MGLC_E_INVALID; }
    MGLC_TRY(require_gpu());
    MGLC_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    mglc_comm *m = new mglc_comm{nullptr, nranks, rank, device};
    ncclResult_t r = ncclCommInitRank(&m->nccl, nranks, uid, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", ncclGetErrorString(r)); delete m; return MGLC_E_NCCL; }
    *c = m;
    return MGLC_OK;
}
extern "C" int mglc_comm_destroy(mglc_comm *c) {
    if (!c) return MGLC_OK;
    cudaSetDevice(c->device);
    ncclCommDestroy(c->nccl);
    delete c;
    return MGLC_OK;
}

// ---- create / destroy -------------------------------------------------------------------------------
static int validate(const mglc_lbm_desc *d) {
    if (d->lattice != MGLC_D3Q19) { set_error("mglc_lbm_create: lattice %d not supported by this handle type", d->lattice); return MGLC_E_INVALID; }
    if (d->collision != MGLC_MRT_LID) { set_error("mglc_lbm_create: collision operator %d not supported", d->collision); return MGLC_E_INVALID; }
    for (int q = 0; q < 3; ++q) {
        if (d->ln[q] < 1 || d->gn[q] < d->ln[q] || d->dims[q] < 1 || d->coords[q] < 0 || d->coords[q] >= d->dims[q] ||
            d->start[q] < 0 || d->start[q] + d->ln[q] > d->gn[q]) {
            set_error("mglc_lbm_create: inconsistent decomposition along dim %d (gn=%d ln=%d start=%d dims=%d coords=%d)", q,
                      d->gn[q], d->ln[q], d->start[q], d->dims[q], d->coords[q]);
            return MGLC_E_INVALID;
        }
    }
    if (!(d->tau > 0.5)) { set_error("mglc_lbm_create: tau=%g must exceed 0.5", d->tau); return MGLC_E_INVALID; }
    return MGLC_OK;
}

extern "C" int mglc_lbm_destroy(mglc_lbm *h) {
    if (!h) return MGLC_OK;
    cudaSetDevice(h->d.device);
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) cudaFree(h->buf[b]);
    double *fields[] = {h->rho, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->stage, h->rho_lid[0], h->rho_lid[1]};
    for (double *p : fields) cudaFree(p);
    for (int m = 0; m < h->nmsgs; ++m) { cudaFree(h->msgs[m].sbuf); cudaFree(h->msgs[m].rbuf); }
    cudaEvent_t evs[] = {h->ev_packed, h->ev_copied, h->ev_t0, h->ev_t1, h->ev_shell, h->ev_halo};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (h->prof_ev) { for (cudaEvent_t e : *h->prof_ev) cudaEventDestroy(e); delete h->prof_ev; }
    if (h->s) cudaStreamDestroy(h->s);
    if (h->s_comm) cudaStreamDestroy(h->s_comm);
    delete h;
    return MGLC_OK;
}

static int create_impl(mglc_lbm **out, const mglc_lbm_desc *d, mglc_comm *comm) {
    if (!out || !d) { set_error("mglc_lbm_create: null argument"); return MGLC_E_INVALID; }
    MGLC_TRY(validate(d));
    MGLC_TRY(require_gpu());
    MGLC_CUDA(cudaSetDevice(d->device));
    mglc_lbm *h = new mglc_lbm();
    memset(h, 0, sizeof *h);
    h->d = *d;
    h->g = make_geom(d->ln[0], d->ln[1], d->ln[2]);
    for (int q = 0; q < 3; ++q) {
        h->g.wall[2 * q] = (d->coords[q] == d->dims[q] - 1);      // +face is a physical wall
        h->g.wall[2 * q + 1] = (d->coords[q] == 0);               // -face
    }
    h->g.lid = h->g.wall[4];                                      // coords(2) == dims(2)-1, L3/bounce_back.f90:72
    mglc_relaxation_rates(d->tau, &h->p.Snu, &h->p.Sq);
    h->p.U0 = d->U0; h->p.rho0 = d->rho0;
    h->nranks = d->dims[0] * d->dims[1] * d->dims[2];
    mglc_cart_rank(d->dims, d->coords, &h->rank);
    h->comm = comm;
    int rc = MGLC_OK;
    auto fail = [&](int code) { mglc_lbm_destroy(h); return code; };
    if ((rc = cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking) == cudaSuccess ? MGLC_OK : MGLC_E_CUDA)) return fail(rc);
    if ((rc = cudaStreamCreateWithFlags(&h->s_comm, cudaStreamNonBlocking) == cudaSuccess ? MGLC_OK : MGLC_E_CUDA)) return fail(rc);
    cudaEvent_t *evs[] = {&h->ev_packed, &h->ev_copied, &h->ev_shell, &h->ev_halo};
    for (cudaEvent_t *e : evs) if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess) return fail(MGLC_E_CUDA);
    const long long nlat = Q * h->g.sq;
    for (int b = 0; b < 2; ++b) {
        if ((rc = dmalloc(h, &h->buf[b], nlat))) return fail(rc);
        if (cudaMemsetAsync(h->buf[b], 0, (size_t)nlat * sizeof(double), h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    }
    const long long n = ncell(h);
    double **fields[] = {&h->rho, &h->u, &h->v, &h->w};
    for (double **f : fields) {
        if ((rc = dmalloc(h, f, n))) return fail(rc);
        if (cudaMemsetAsync(*f, 0, (size_t)n * sizeof(double), h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    }
    if ((rc = dmalloc(h, &h->scratch, check_scratch_doubles()))) return fail(rc);
    for (int b = 0; b < 2; ++b)
        if ((rc = dmalloc(h, &h->rho_lid[b], (long long)h->g.nx * h->g.ny))) return fail(rc);
    h->lid_next = h->lid_last_in = h->rho + (long long)h->g.nx * h->g.ny * (h->g.nz - 1);
    // halo plan + buffers
    mglc_halo_msg plan[18];
    mglc_halo_plan(d, plan, &h->nmsgs);
    for (int m = 0; m < h->nmsgs; ++m) {
        Msg &M = h->msgs[m];
        M.dir = plan[m].dir; M.send_to = plan[m].send_to; M.recv_from = plan[m].recv_from;
        M.send_count = plan[m].send_count; M.recv_count = plan[m].recv_count;
        M.sbuf = M.rbuf = nullptr;
        if (M.send_count && (rc = dmalloc(h, &M.sbuf, M.send_count))) return fail(rc);
        if (M.recv_count && (rc = dmalloc(h, &M.rbuf, M.recv_count))) return fail(rc);
        if (M.send_count || M.recv_count) h->has_neighbors = true;
    }
    if (cudaStreamSynchronize(h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_lbm_create(mglc_lbm **h, const mglc_lbm_desc *d, mglc_comm *comm_or_null) {
    if (d && d->dims[0] * d->dims[1] * d->dims[2] > 1 && !comm_or_null) {
        set_error("mglc_lbm_create: a %dx%dx%d decomposition needs a communicator (or use mglc_group_create)",
                  d->dims[0], d->dims[1], d->dims[2]);
        return MGLC_E_INVALID;
    }
    return create_impl(h, d, comm_or_null);
}

extern "C" int mglc_lbm_get_desc(mglc_lbm *h, mglc_lbm_desc *d) {
    if (!h || !d) return MGLC_E_INVALID;
    *d = h->d;
    return MGLC_OK;
}
extern "C" int mglc_lbm_device_bytes(mglc_lbm *h, long long *bytes) {
    if (!h || !bytes) return MGLC_E_INVALID;
    *bytes = h->bytes;
    return MGLC_OK;
}
extern "C" int mglc_lbm_launch_count(mglc_lbm *h, long long *n) {
    if (!h || !n) return MGLC_E_INVALID;
    *n = h->launches;
    return MGLC_OK;
}
extern "C" int mglc_lbm_sync(mglc_lbm *h) {
    MGLC_TRY(use(h));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s_comm));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

static int canonicalise(mglc_lbm *h);

// ---- host <-> device transfers in the reference layout ---------------------------------------------------
static int ensure_stage(mglc_lbm *h) {
    if (h->stage) return MGLC_OK;
    h->stage_doubles = 8LL << 20;   // 64 MiB
    return dmalloc(h, &h->stage, h->stage_doubles);
}
// f (0:18,nx,ny,nz) or f_post (0:18,0:nx+1,...) on the host <-> SoA lattice on the device
static int transfer_lattice(mglc_lbm *h, double *host, double *dev, int with_halo, bool to_device) {
    MGLC_TRY(ensure_stage(h));
    const int e = with_halo ? 2 : 0;
    const long long total = (long long)(h->g.nx + e) * (h->g.ny + e) * (h->g.nz + e);
    const long long chunk = h->stage_doubles / Q;
    for (long long c0 = 0; c0 < total; c0 += chunk) {
        const long long nc = std::min(chunk, total - c0);
        if (to_device) {
            MGLC_CUDA(cudaMemcpyAsync(h->stage, host + c0 * Q, (size_t)nc * Q * sizeof(double), cudaMemcpyHostToDevice, h->s));
            h->launches += launch_aos_to_soa(h->g, h->stage, dev, c0, nc, with_halo, h->s);
        } else {
            h->launches += launch_soa_to_aos(h->g, dev, h->stage, c0, nc, with_halo, h->s);
            MGLC_CUDA(cudaMemcpyAsync(host + c0 * Q, h->stage, (size_t)nc * Q * sizeof(double), cudaMemcpyDeviceToHost, h->s));
        }
    }
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return MGLC_OK;
}
static int copy_field(mglc_lbm *h, double *host, double *dev, bool to_device) {
    if (!host) return MGLC_OK;
    const size_t bytes = (size_t)ncell(h) * sizeof(double);
    if (to_device) MGLC_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, h->s));
    else MGLC_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->s));
    return MGLC_OK;
}

extern "C" int mglc_lbm_upload(mglc_lbm *h, const double *f, const double *rho, const double *u, const double *v,
                               const double *w) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (f) MGLC_TRY(transfer_lattice(h, const_cast<double *>(f), F_(h), 0, true));
    MGLC_TRY(copy_field(h, const_cast<double *>(rho), h->rho, true));
    MGLC_TRY(copy_field(h, const_cast<double *>(u), h->u, true));
    MGLC_TRY(copy_field(h, const_cast<double *>(v), h->v, true));
    MGLC_TRY(copy_field(h, const_cast<double *>(w), h->w, true));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return MGLC_OK;
}
extern "C" int mglc_lbm_upload_fpost(mglc_lbm *h, const double *f_post) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (!f_post) return MGLC_E_INVALID;
    return transfer_lattice(h, const_cast<double *>(f_post), Fpost_(h), 1, true);
}
extern "C" int mglc_lbm_download_macro(mglc_lbm *h, double *rho, double *u, double *v, double *w) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    MGLC_TRY(copy_field(h, rho, h->rho, false));
    MGLC_TRY(copy_field(h, u, h->u, false));
    MGLC_TRY(copy_field(h, v, h->v, false));
    MGLC_TRY(copy_field(h, w, h->w, false));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return MGLC_OK;
}
extern "C" int mglc_lbm_download_f(mglc_lbm *h, double *f) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (!f) return MGLC_E_INVALID;
    return transfer_lattice(h, f, F_(h), 0, false);
}
extern "C" int mglc_lbm_download_fpost(mglc_lbm *h, double *f_post) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (!f_post) return MGLC_E_INVALID;
    return transfer_lattice(h, f_post, Fpost_(h), 1, false);
}

// ---- single-subdomain building blocks ---------------------------------------------------------------------
static int do_initial(mglc_lbm *h) {
    h->rotated = 0;
    h->lid_next = h->lid_last_in = h->rho + (long long)h->g.nx * h->g.ny * (h->g.nz - 1);
    h->launches += launch_initial(h->g, h->p, F_(h), h->rho, h->u, h->v, h->w, h->s);
    if (h->up) {
        const size_t b = (size_t)ncell(h) * sizeof(double);
        MGLC_CUDA(cudaMemsetAsync(h->up, 0, b, h->s));
        MGLC_CUDA(cudaMemsetAsync(h->vp, 0, b, h->s));
        MGLC_CUDA(cudaMemsetAsync(h->wp, 0, b, h->s));
    }
    return MGLC_OK;
}
static int do_collision(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    h->launches += strict_(h) ? strict::launch_collision(h->g, h->p, F_(h), h->rho, h->u, h->v, h->w, Fpost_(h), h->s)
                              : fast::launch_collision(h->g, h->p, F_(h), h->rho, h->u, h->v, h->w, Fpost_(h), h->s);
    return MGLC_OK;
}
static int do_pack(mglc_lbm *h, cudaStream_t s) {
    for (int m = 0; m < h->nmsgs; ++m)
        if (h->msgs[m].send_count) h->launches += launch_pack(h->g, Fpost_(h), h->msgs[m].dir, h->msgs[m].sbuf, s);
    return MGLC_OK;
}
static int do_unpack(mglc_lbm *h, cudaStream_t s) {
    for (int m = 0; m < h->nmsgs; ++m)
        if (h->msgs[m].recv_count) h->launches += launch_unpack(h->g, Fpost_(h), h->msgs[m].dir, h->msgs[m].rbuf, s);
    return MGLC_OK;
}
static int do_nccl_sendrecv(mglc_lbm *h, cudaStream_t s) {
    MGLC_NCCL(ncclGroupStart());
    for (int m = 0; m < h->nmsgs; ++m) {
        Msg &M = h->msgs[m];
        if (M.send_count) MGLC_NCCL(ncclSend(M.sbuf, (size_t)M.send_count, ncclDouble, M.send_to, h->comm->nccl, s));
        if (M.recv_count) MGLC_NCCL(ncclRecv(M.rbuf, (size_t)M.recv_count, ncclDouble, M.recv_from, h->comm->nccl, s));
    }
    MGLC_NCCL(ncclGroupEnd());
    return MGLC_OK;
}
// message_passing_sendrecv() for a handle that owns a communicator (one process per GPU)
static int do_exchange_nccl(mglc_lbm *h) {
    if (!h->has_neighbors) return MGLC_OK;
    if (!h->comm) { set_error("exchange: subdomain has neighbours but no communicator"); return MGLC_E_STATE; }
    MGLC_TRY(do_pack(h, h->s));
    MGLC_TRY(do_nccl_sendrecv(h, h->s));
    MGLC_TRY(do_unpack(h, h->s));
    return MGLC_OK;
}
static int do_streaming(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    h->launches += launch_streaming(h->g, Fpost_(h), F_(h), h->s);
    return MGLC_OK;
}
static int do_bounceback(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    h->launches += launch_bounceback(h->g, h->p, Fpost_(h), h->rho, F_(h), h->s);
    return MGLC_OK;
}
static int do_macro(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    h->launches += launch_macro(h->g, F_(h), h->rho, h->u, h->v, h->w, h->s);
    return MGLC_OK;
}
// fused stream+macro+collide over the whole subdomain: reads f_post (incl. halo), writes the NEXT f_post
// into the buffer that held f, then the two buffers swap roles
constexpr int PROF_PAIRS = 64;
static int prof_flush(mglc_lbm *h) {
    if (!h->prof_ev || !h->prof_used) return MGLC_OK;
    MGLC_CUDA(cudaEventSynchronize((*h->prof_ev)[2 * (h->prof_used - 1) + 1]));
    for (int q = 0; q < h->prof_used; ++q) {
        float t = 0.f;
        MGLC_CUDA(cudaEventElapsedTime(&t, (*h->prof_ev)[2 * q], (*h->prof_ev)[2 * q + 1]));
        h->prof_ms += t;
        h->prof_launches += 1;
    }
    h->prof_used = 0;
    return MGLC_OK;
}
static int do_fused(mglc_lbm *h) {
    const int box[6] = {1, h->g.nx, 1, h->g.ny, 1, h->g.nz};
    if (h->profiling) {
        if (!h->prof_ev) {
            h->prof_ev = new std::vector<cudaEvent_t>(2 * PROF_PAIRS);
            for (cudaEvent_t &e : *h->prof_ev) MGLC_CUDA(cudaEventCreate(&e));
        }
        if (h->prof_used == PROF_PAIRS) MGLC_TRY(prof_flush(h));
        MGLC_CUDA(cudaEventRecord((*h->prof_ev)[2 * h->prof_used], h->s));
    }
    const double *lid_in = h->lid_next;
    double *lid_out = (lid_in == h->rho_lid[0]) ? h->rho_lid[1] : h->rho_lid[0];
    h->launches += strict_(h) ? strict::launch_fused(h->g, h->p, Fpost_(h), F_(h), lid_in, lid_out, box, h->s)
                              : fast::launch_fused(h->g, h->p, Fpost_(h), F_(h), lid_in, lid_out, box, h->s);
    h->lid_last_in = lid_in;
    h->lid_next = lid_out;
    if (h->profiling) {
        MGLC_CUDA(cudaEventRecord((*h->prof_ev)[2 * h->prof_used + 1], h->s));
        h->prof_used += 1;
    }
    h->cur ^= 1;
    return MGLC_OK;
}
static int do_stream_macro(mglc_lbm *h) {
    h->launches += strict_(h) ? strict::launch_stream_macro(h->g, h->p, Fpost_(h), F_(h), h->lid_last_in, h->rho, h->u, h->v, h->w, h->s)
                              : fast::launch_stream_macro(h->g, h->p, Fpost_(h), F_(h), h->lid_last_in, h->rho, h->u, h->v, h->w, h->s);
    h->lid_next = h->lid_last_in = h->rho + (long long)h->g.nx * h->g.ny * (h->g.nz - 1);
    return MGLC_OK;
}
// Leave the rotated state: the last step's f_post (with halos) is still intact in buf[cur]; pull it into
// the other lattice (discarding the pre-computed next collision, which collision() will redo) and write
// rho,u,v,w.  Afterwards f, f_post, rho,u,v,w are the reference's after the same number of loop bodies.
static int canonicalise(mglc_lbm *h) {
    if (!h->rotated) return MGLC_OK;
    h->cur ^= 1;
    MGLC_TRY(do_stream_macro(h));
    h->rotated = 0;
    return MGLC_OK;
}
static int ensure_prev(mglc_lbm *h) {
    if (h->up) return MGLC_OK;
    const long long n = ncell(h);
    MGLC_TRY(dmalloc(h, &h->up, n));
    MGLC_TRY(dmalloc(h, &h->vp, n));
    MGLC_TRY(dmalloc(h, &h->wp, n));
    MGLC_CUDA(cudaMemsetAsync(h->up, 0, (size_t)n * sizeof(double), h->s));     // up = vp = wp = 0, L3/initial.f90:50-52
    MGLC_CUDA(cudaMemsetAsync(h->vp, 0, (size_t)n * sizeof(double), h->s));
    MGLC_CUDA(cudaMemsetAsync(h->wp, 0, (size_t)n * sizeof(double), h->s));
    return MGLC_OK;
}
static int do_check_partial(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    MGLC_TRY(ensure_prev(h));
    h->launches += launch_check(h->g, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->s);
    return MGLC_OK;
}

static int not_in_group(mglc_lbm *h, const char *what) {
    if (h->group) { set_error("%s: this subdomain belongs to a group; call the mglc_group_* entry point", what); return MGLC_E_STATE; }
    return MGLC_OK;
}

// ---- public per-subroutine entry points ------------------------------------------------------------------
extern "C" int mglc_lbm_initial(mglc_lbm *h) { MGLC_TRY(use(h)); return do_initial(h); }
extern "C" int mglc_collision(mglc_lbm *h) { MGLC_TRY(use(h)); return do_collision(h); }
extern "C" int mglc_exchange(mglc_lbm *h) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_exchange"));
    MGLC_TRY(canonicalise(h));
    return do_exchange_nccl(h);
}
extern "C" int mglc_streaming(mglc_lbm *h) { MGLC_TRY(use(h)); return do_streaming(h); }
extern "C" int mglc_bounceback(mglc_lbm *h) { MGLC_TRY(use(h)); return do_bounceback(h); }
extern "C" int mglc_macro(mglc_lbm *h) { MGLC_TRY(use(h)); return do_macro(h); }

extern "C" int mglc_check(mglc_lbm *h, double *errorU) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_check"));
    if (!errorU) return MGLC_E_INVALID;
    MGLC_TRY(do_check_partial(h));
    if (h->comm && h->nranks > 1)   // MPI_Allreduce(SUM) x2, L3/check.f90:27-28
        MGLC_NCCL(ncclAllReduce(h->scratch, h->scratch, 2, ncclDouble, ncclSum, h->comm->nccl, h->s));
    double e[2];
    MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    *errorU = sqrt(e[0]) / sqrt(e[1]);
    return MGLC_OK;
}

// nsteps iterations of: collision, exchange, streaming, bounceback, macro  (L3/main.f90:85-97).
// Rotated by half a step: [collision once] then nsteps x (exchange -> fused pull+walls+macro+collide).
// The handle stays rotated afterwards, so back-to-back calls pay neither prologue nor epilogue; any
// entry point that needs the reference's state calls canonicalise() first.
static int step_impl(mglc_lbm *h, int nsteps) {
    if (nsteps < 0) { set_error("mglc_lbm_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    if (nsteps == 0) return MGLC_OK;
    if (!h->rotated) MGLC_TRY(do_collision(h));      // the first step's collision()
    for (int it = 0; it < nsteps; ++it) {
        MGLC_TRY(do_exchange_nccl(h));               // exchange of step it
        MGLC_TRY(do_fused(h));                       // streaming+bounceback+macro of step it, collision of it+1
    }
    h->rotated = 1;
    return MGLC_OK;
}
extern "C" int mglc_lbm_step(mglc_lbm *h, int nsteps) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_lbm_step"));
    MGLC_TRY(step_impl(h, nsteps));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}
extern "C" int mglc_lbm_step_timed(mglc_lbm *h, int nsteps, float *ms) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_lbm_step_timed"));
    if (!ms) return MGLC_E_INVALID;
    MGLC_CUDA(cudaEventRecord(h->ev_t0, h->s));
    MGLC_TRY(step_impl(h, nsteps));
    MGLC_CUDA(cudaEventRecord(h->ev_t1, h->s));
    MGLC_CUDA(cudaEventSynchronize(h->ev_t1));
    MGLC_CUDA(cudaGetLastError());
    MGLC_CUDA(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
    return MGLC_OK;
}
extern "C" int mglc_lbm_set_profiling(mglc_lbm *h, int on) {
    MGLC_TRY(use(h));
    MGLC_TRY(prof_flush(h));
    h->profiling = on ? 1 : 0;
    return MGLC_OK;
}
extern "C" int mglc_lbm_kernel_time(mglc_lbm *h, float *fused_ms, long long *fused_launches) {
    MGLC_TRY(use(h));
    MGLC_TRY(prof_flush(h));
    if (fused_ms) *fused_ms = (float)h->prof_ms;
    if (fused_launches) *fused_launches = h->prof_launches;
    h->prof_ms = 0.0;
    h->prof_launches = 0;
    return MGLC_OK;
}
// pinned host memory for the caller's arrays (a Fortran driver maps it with c_f_pointer)
extern "C" int mglc_host_alloc(void **p, size_t bytes) {
    if (!p) return MGLC_E_INVALID;
    MGLC_TRY(require_gpu());
    MGLC_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return MGLC_OK;
}
extern "C" int mglc_host_free(void *p) {
    if (p) MGLC_CUDA(cudaFreeHost(p));
    return MGLC_OK;
}

// ---- P subdomains in one process ----------------------------------------------------------------------------
extern "C" int mglc_group_destroy(mglc_group *g) {
    if (!g) return MGLC_OK;
    for (mglc_lbm *h : g->r) mglc_lbm_destroy(h);
    delete g;
    return MGLC_OK;
}
extern "C" int mglc_group_create(mglc_group **out, const mglc_lbm_desc *gd, int nranks, const int *devices) {
    if (!out || !gd || nranks < 1) { set_error("mglc_group_create: bad arguments"); return MGLC_E_INVALID; }
    MGLC_TRY(require_gpu());
    mglc_group *g = new mglc_group();
    g->global = *gd;
    for (int r = 0; r < nranks; ++r) {
        mglc_lbm_desc d = *gd;
        if (d.dims[0] <= 0) mglc_dims_create(nranks, d.dims);
        if (d.dims[0] * d.dims[1] * d.dims[2] != nranks) { set_error("mglc_group_create: dims do not multiply to nranks"); mglc_group_destroy(g); return MGLC_E_INVALID; }
        mglc_cart_coords(d.dims, r, d.coords);
        for (int q = 0; q < 3; ++q) {
            int rc = mglc_decompose_1d(d.gn[q], d.coords[q], d.dims[q], &d.ln[q], &d.start[q]);
            if (rc) { mglc_group_destroy(g); return rc; }
        }
        d.device = devices ? devices[r] : gd->device;
        mglc_lbm *h = nullptr;
        int rc = create_impl(&h, &d, nullptr);
        if (rc) { mglc_group_destroy(g); return rc; }
        h->group = g;
        g->r.push_back(h);
    }
    // enable peer access where subdomains live on different devices (best effort; copies fall back to staging)
    for (mglc_lbm *a : g->r)
        for (mglc_lbm *b : g->r)
            if (a->d.device != b->d.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->d.device, b->d.device);
                if (can) { cudaSetDevice(a->d.device); cudaDeviceEnablePeerAccess(b->d.device, 0); (void)cudaGetLastError(); }
            }
    *out = g;
    return MGLC_OK;
}
extern "C" int mglc_group_size(mglc_group *g, int *n) { if (!g || !n) return MGLC_E_INVALID; *n = (int)g->r.size(); return MGLC_OK; }
extern "C" int mglc_group_rank(mglc_group *g, int r, mglc_lbm **h) {
    if (!g || !h || r < 0 || r >= (int)g->r.size()) return MGLC_E_INVALID;
    *h = g->r[r];
    return MGLC_OK;
}

#define FOR_RANKS(g, h) for (mglc_lbm * h : (g)->r)
#define GROUP_EACH(g, fn)                       \
    do {                                        \
        if (!(g)) return MGLC_E_INVALID;        \
        FOR_RANKS(g, h_) { MGLC_TRY(use(h_)); MGLC_TRY(fn(h_)); } \
        return MGLC_OK;                         \
    } while (0)

// message_passing_sendrecv() across the group: pack on every sender, receiver-driven device-to-device
// copies (ordered by events), unpack on every receiver
static int group_exchange(mglc_group *g) {
    FOR_RANKS(g, h) {
        if (!h->has_neighbors) continue;
        MGLC_TRY(use(h));
        // the peers that copied out of my send buffers last time must be done before I overwrite them
        for (int m = 0; m < h->nmsgs; ++m)
            if (h->msgs[m].send_count) MGLC_CUDA(cudaStreamWaitEvent(h->s, g->r[h->msgs[m].send_to]->ev_copied, 0));
        MGLC_TRY(do_pack(h, h->s));
        MGLC_CUDA(cudaEventRecord(h->ev_packed, h->s));
    }
    FOR_RANKS(g, h) {
        if (!h->has_neighbors) continue;
        MGLC_TRY(use(h));
        for (int m = 0; m < h->nmsgs; ++m) {
            Msg &M = h->msgs[m];
            if (!M.recv_count) continue;
            mglc_lbm *src = g->r[M.recv_from];
            MGLC_CUDA(cudaStreamWaitEvent(h->s, src->ev_packed, 0));
            MGLC_CUDA(cudaMemcpyPeerAsync(M.rbuf, h->d.device, src->msgs[m].sbuf, src->d.device,
                                          (size_t)M.recv_count * sizeof(double), h->s));
        }
        MGLC_CUDA(cudaEventRecord(h->ev_copied, h->s));
        MGLC_TRY(do_unpack(h, h->s));
    }
    return MGLC_OK;
}

extern "C" int mglc_group_initial(mglc_group *g) { GROUP_EACH(g, do_initial); }
extern "C" int mglc_group_collision(mglc_group *g) { GROUP_EACH(g, do_collision); }
extern "C" int mglc_group_exchange(mglc_group *g) {
    if (!g) return MGLC_E_INVALID;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(canonicalise(h)); }
    return group_exchange(g);
}
extern "C" int mglc_group_streaming(mglc_group *g) { GROUP_EACH(g, do_streaming); }
extern "C" int mglc_group_bounceback(mglc_group *g) { GROUP_EACH(g, do_bounceback); }
extern "C" int mglc_group_macro(mglc_group *g) { GROUP_EACH(g, do_macro); }

extern "C" int mglc_group_check(mglc_group *g, double *errorU) {
    if (!g || !errorU) return MGLC_E_INVALID;
    double t1 = 0.0, t2 = 0.0;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(do_check_partial(h)); }
    FOR_RANKS(g, h) {                       // Allreduce(SUM) modelled as a rank-ordered host sum
        MGLC_TRY(use(h));
        double e[2];
        MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
        MGLC_CUDA(cudaStreamSynchronize(h->s));
        t1 += e[0]; t2 += e[1];
    }
    *errorU = sqrt(t1) / sqrt(t2);
    return MGLC_OK;
}

static int group_step_impl(mglc_group *g, int nsteps) {
    if (nsteps < 0) return MGLC_E_INVALID;
    if (nsteps == 0) return MGLC_OK;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); if (!h->rotated) MGLC_TRY(do_collision(h)); }
    for (int it = 0; it < nsteps; ++it) {
        MGLC_TRY(group_exchange(g));
        FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(do_fused(h)); }
    }
    FOR_RANKS(g, h) h->rotated = 1;
    return MGLC_OK;
}
extern "C" int mglc_group_step(mglc_group *g, int nsteps) {
    if (!g) return MGLC_E_INVALID;
    MGLC_TRY(group_step_impl(g, nsteps));
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
extern "C" int mglc_group_step_timed(mglc_group *g, int nsteps, float *ms) {
    if (!g || !ms) return MGLC_E_INVALID;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_CUDA(cudaStreamSynchronize(h->s)); }
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_CUDA(cudaEventRecord(h->ev_t0, h->s)); }
    MGLC_TRY(group_step_impl(g, nsteps));
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_CUDA(cudaEventRecord(h->ev_t1, h->s)); }
    float worst = 0.f;
    FOR_RANKS(g, h) {
        MGLC_TRY(use(h));
        MGLC_CUDA(cudaEventSynchronize(h->ev_t1));
        float t = 0.f;
        MGLC_CUDA(cudaEventElapsedTime(&t, h->ev_t0, h->ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MGLC_OK;
}
