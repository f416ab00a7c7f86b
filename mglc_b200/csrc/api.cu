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
    // overlapped schedule (L3nb:1108-1230): the halo exchange of the lattice just written runs on s_comm while the
    // interior update runs on s.  halo_inflight: that exchange has been issued for buf[cur^1]; s must wait for
    // ev_halo before anything reads those halos.  overlap: 1 = use the schedule when the handle has neighbours + NCCL
    int halo_inflight, overlap;
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
    cudaEvent_t ev_stage_full[2], ev_stage_free[2];      // double-buffered host<->device staging (transfer_lattice)
    long long launches;
    long long bytes;
    Msg msgs[24];            // 0..17 f messages (faces, edges); 18..23 g faces (dir 20+face) for thermal handles
    int nmsgs;
    // thermal double-distribution state (lattice == MGLC_D3Q19_D3Q7)
    bool thermal;
    ThermalParams tp;
    double *gbuf[2];         // gbuf[cur] = g, gbuf[cur^1] = g_post (ping-pong in lockstep with buf)
    double *T, *Tp;
    double *Fc[2];           // Fx,Fy,Fz back to back; Fc[fc_cur] = the force collision() wrote / macro() reads
    int fc_cur;
    bool has_neighbors;
    mglc_comm *comm;
    mglc_group *group;
    // direct halo stores (PeerTable, common.cuh): overlap == 2.  pt_dev[b] is the table for a launch that writes
    // buf[b]; flags = the 32 barrier words neighbours raise; epoch counts direct launches; direct_valid: the halos
    // of the lattice written last are being filled by the neighbours' launches of the same epoch
    int direct, direct_valid;
    PeerTable *pt_dev[2];
    SyncTable sync;
    unsigned long long *flags, epoch;
    int parity_sent;                     // ping-pong index of the lattice my last direct launch wrote (neighbours must agree)
    int *d_err;
    // halo push: staging of the x-face messages I RECEIVE, [lattice parity][message 0 / 1][xs_count doubles] (PeerTable::XS);
    // xs_pending: my neighbours' pushes of the current epoch went there and still have to be scattered into my halo columns
    double *xstage;
    long long xs_count;
    int xs_pending;
    int nbr_rank[19];                    // rank at coords + e_d, -1 = none
    cudaEvent_t ev_done[2];              // group mode: my direct launch of epoch e is recorded in ev_done[e & 1]
    std::vector<void *> *ipc_opened;     // comm mode: mappings to close on destroy
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
    std::vector<Port> ports;
    mglc_lbm_desc global;
};
enum { MSG_F = 1, MSG_G = 2, MSG_ALL = 3 };

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
static inline double *G_(mglc_lbm *h) { return h->gbuf[h->cur]; }
static inline double *Gpost_(mglc_lbm *h) { return h->gbuf[h->cur ^ 1]; }
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
    if (d->lattice != MGLC_D3Q19 && d->lattice != MGLC_D3Q19_D3Q7) { set_error("mglc_lbm_create: lattice %d not supported by this handle type", d->lattice); return MGLC_E_INVALID; }
    if (d->collision != (d->lattice == MGLC_D3Q19 ? MGLC_MRT_LID : MGLC_MRT_THERMAL) &&
        !(d->lattice == MGLC_D3Q19 && d->collision == MGLC_BGK)) { set_error("mglc_lbm_create: collision operator %d not supported with lattice %d", d->collision, d->lattice); return MGLC_E_INVALID; }
    if (d->lattice == MGLC_D3Q19_D3Q7) {
        for (int f = 0; f < 6; ++f)
            if (d->bcT[f] < MGLC_BCT_ADIABATIC || d->bcT[f] > MGLC_BCT_CONST_COLD) { set_error("mglc_lbm_create: bcT[%d]=%d", f, d->bcT[f]); return MGLC_E_INVALID; }
        if (d->bcT[0] || d->bcT[1]) { set_error("mglc_lbm_create: the reference has no constant-temperature x walls (B3:1186-1205)"); return MGLC_E_INVALID; }
    }
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

static int wait_direct(mglc_lbm *h);
extern "C" int mglc_lbm_destroy(mglc_lbm *h) {
    if (!h) return MGLC_OK;
    cudaSetDevice(h->d.device);
    if (h->direct && h->direct_valid && h->s) wait_direct(h);      // neighbours may still be storing into my halos
    cudaDeviceSynchronize();
    if (h->ipc_opened) { for (void *q : *h->ipc_opened) cudaIpcCloseMemHandle(q); delete h->ipc_opened; }
    cudaFree(h->pt_dev[0]); cudaFree(h->pt_dev[1]); cudaFree(h->flags); cudaFree(h->d_err); cudaFree(h->xstage);
    for (cudaEvent_t e : h->ev_done) if (e) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) cudaFree(h->buf[b]);
    double *fields[] = {h->rho, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->stage, h->rho_lid[0], h->rho_lid[1],
                        h->gbuf[0], h->gbuf[1], h->T, h->Tp, h->Fc[0], h->Fc[1]};
    for (double *p : fields) cudaFree(p);
    for (int m = 0; m < h->nmsgs; ++m) { cudaFree(h->msgs[m].sbuf); cudaFree(h->msgs[m].rbuf); }
    cudaEvent_t evs[] = {h->ev_packed, h->ev_copied, h->ev_t0, h->ev_t1, h->ev_shell, h->ev_halo,
                         h->ev_stage_full[0], h->ev_stage_full[1], h->ev_stage_free[0], h->ev_stage_free[1]};
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
    h->thermal = (d->lattice == MGLC_D3Q19_D3Q7);
    h->g.lid = h->thermal ? 0 : h->g.wall[4];                     // coords(2) == dims(2)-1, L3/bounce_back.f90:72
    mglc_relaxation_rates(d->tau, &h->p.Snu, &h->p.Sq);
    h->p.U0 = d->U0; h->p.rho0 = d->rho0;
    h->p.bgk = d->collision == MGLC_BGK;
    if (h->thermal) {
        ThermalParams &tp = h->tp;
        tp.Snu = h->p.Snu; tp.Sq = h->p.Sq; tp.Qd = d->Qd; tp.Qnu = d->Qnu; tp.paraA = d->paraA; tp.gBeta = d->gBeta;
        tp.Tref = d->Tref; tp.omegaRot = d->omegaRot; tp.Thot = d->Thot; tp.Tcold = d->Tcold;
        for (int f = 0; f < 6; ++f) {
            tp.bcT[f] = d->bcT[f];
            const double Tw = d->bcT[f] == MGLC_BCT_CONST_HOT ? d->Thot : d->Tcold;
            tp.wallT[f] = (6.0 + d->paraA) / 21.0 * Tw;              // B3:1128,1137,1148,1157
        }
    }
    h->nranks = d->dims[0] * d->dims[1] * d->dims[2];
    mglc_cart_rank(d->dims, d->coords, &h->rank);
    h->comm = comm;
    h->overlap = 1;
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
    if (h->thermal) {
        for (int b = 0; b < 2; ++b) {
            if ((rc = dmalloc(h, &h->gbuf[b], QT * h->g.sq))) return fail(rc);
            if (cudaMemsetAsync(h->gbuf[b], 0, (size_t)QT * h->g.sq * sizeof(double), h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
            if ((rc = dmalloc(h, &h->Fc[b], 3 * n))) return fail(rc);
            if (cudaMemsetAsync(h->Fc[b], 0, (size_t)3 * n * sizeof(double), h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
        }
        if ((rc = dmalloc(h, &h->T, n))) return fail(rc);
        if (cudaMemsetAsync(h->T, 0, (size_t)n * sizeof(double), h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
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
    if (h->thermal) {              // g_message_passing_sendrecv(): one population per face, B3:1421-1468
        for (int face = 0; face < 6; ++face) {
            Msg &M = h->msgs[h->nmsgs++];
            const Msg &Ff = h->msgs[face];                        // same neighbours and slab as f's face message
            M.dir = 20 + face; M.send_to = Ff.send_to; M.recv_from = Ff.recv_from; M.skip = 0;
            M.send_count = Ff.send_count / 5; M.recv_count = Ff.recv_count / 5;
            M.sbuf = M.rbuf = nullptr;
            if (M.send_count && (rc = dmalloc(h, &M.sbuf, M.send_count))) return fail(rc);
            if (M.recv_count && (rc = dmalloc(h, &M.rbuf, M.recv_count))) return fail(rc);
        }
    }
    // neighbour table and the barrier words of the direct-halo path (set up by setup_direct_* once peers are known)
    for (int dd = 0; dd < 19; ++dd) {
        h->nbr_rank[dd] = -1;
        if (dd == 6) continue;
        const int e[3] = {dd < 6 ? (dd >> 1 == 0 ? 1 - 2 * (dd & 1) : 0) : h_ex[dd], dd < 6 ? (dd >> 1 == 1 ? 1 - 2 * (dd & 1) : 0) : h_ey[dd],
                          dd < 6 ? (dd >> 1 == 2 ? 1 - 2 * (dd & 1) : 0) : h_ez[dd]};
        const int c[3] = {d->coords[0] + e[0], d->coords[1] + e[1], d->coords[2] + e[2]};
        mglc_cart_rank(d->dims, c, &h->nbr_rank[dd]);
    }
    if (cudaMalloc((void **)&h->flags, 32 * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc((void **)&h->d_err, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&h->pt_dev[0], sizeof(PeerTable)) != cudaSuccess || cudaMalloc((void **)&h->pt_dev[1], sizeof(PeerTable)) != cudaSuccess)
        return fail(MGLC_E_NOMEM);
    if (h->nbr_rank[0] >= 0 || h->nbr_rank[1] >= 0) {
        h->xs_count = (long long)(h->thermal ? 6 : 5) * h->g.ny * h->g.nz;
        if ((rc = dmalloc(h, &h->xstage, 4 * h->xs_count))) return fail(rc);
    }
    cudaMemsetAsync(h->flags, 0, 32 * sizeof(unsigned long long), h->s);
    cudaMemsetAsync(h->d_err, 0, sizeof(int), h->s);
    for (cudaEvent_t &e : h->ev_done) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaStreamSynchronize(h->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    *out = h;
    return MGLC_OK;
}

// Which of the two NVLink transports a handle starts with once its neighbours are mapped.  Measured on 2 and 8 B200s (768^3
// blocks, profiles/r2g_*, r2m_*): the push after the update costs 0.05 ms per step for a y or z face and 0.24 ms for an x face,
// the stores from inside the update kernel 0.2 and 0.57 ms (8 GPUs, 2x2x2: 21.50 against 22.33 ms); on 256^3 blocks (thermal,
// config 4) the in-kernel stores are 1 % ahead because the two extra launches of the push weigh more than the stores.  Both
// costs scale with the surface, the launches do not: the switch sits between 256^3 and 384^3.  MGLC_HALO_MODE=2|3 overrides.
static int default_direct_mode(const mglc_lbm *h) {
    if (const char *e = getenv("MGLC_HALO_MODE")) { const int v = atoi(e); if (v == 2 || v == 3) return v; }
    return (long long)h->g.nx * h->g.ny * h->g.nz >= (1LL << 25) ? 3 : 2;
}

// ---- direct halo stores: wiring ------------------------------------------------------------------------------
static inline int opp_dir(int d) { return d < 6 ? (d ^ 1) : h_opp[d]; }
struct PeerView { double *buf[2], *gbuf[2], *xstage; unsigned long long *flags; int ln[3]; };
// fill pt_dev[0..1] and sync from the neighbours' lattices as seen from this GPU
static int install_peers(mglc_lbm *h, const PeerView *view /* [19], valid where nbr_rank >= 0 */) {
    PeerTable pt[2];
    memset(pt, 0, sizeof pt);
    memset(&h->sync, 0, sizeof h->sync);
    for (int d = 0; d < 19; ++d) {
        if (d == 6 || h->nbr_rank[d] < 0) continue;
        const Geom pg = make_geom(view[d].ln[0], view[d].ln[1], view[d].ln[2]);
        for (int b = 0; b < 2; ++b) {
            pt[b].mask |= 1u << d;
            pt[b].F[d] = view[d].buf[b];
            if (d < 6) pt[b].G[d] = view[d].gbuf[b];
            pt[b].sy[d] = pg.sy; pt[b].sz[d] = pg.sz; pt[b].sq[d] = pg.sq;
            pt[b].n[d][0] = pg.nx; pt[b].n[d][1] = pg.ny; pt[b].n[d][2] = pg.nz;
            // my +x message lands in the +x neighbour's staging of message 0, my -x message in the -x neighbour's of message 1
            // (an x neighbour has my ny, nz, hence my xs_count)
            if (d < 2 && view[d].xstage) pt[b].XS[d] = view[d].xstage + (b * 2 + d) * h->xs_count;
        }
        h->sync.mask |= 1u << d;
        h->sync.signal[d] = view[d].flags + opp_dir(d);     // the neighbour sees me in the opposite direction
        h->sync.wait[d] = h->flags + d;
    }
    for (int b = 0; b < 2; ++b) {
        pt[b].err = h->d_err;
        MGLC_CUDA(cudaMemcpy(h->pt_dev[b], &pt[b], sizeof(PeerTable), cudaMemcpyHostToDevice));
    }
    h->direct = 1;
    return MGLC_OK;
}
// one process per GPU: exchange CUDA IPC handles of the lattices and the barrier words through the communicator,
// map the neighbours' allocations (peer access over NVLink) and agree collectively on whether the path is usable
struct IpcRecord { cudaIpcMemHandle_t buf[2], gbuf[2], flags, xstage; int ln[3]; int thermal, has_xstage; };
static int setup_direct_ipc(mglc_lbm *h) {
    if (!h->comm || h->nranks < 2 || getenv("MGLC_NO_DIRECT")) return MGLC_OK;
    const int P = h->nranks;
    IpcRecord mine;
    memset(&mine, 0, sizeof mine);
    int ok = 1;
    const unsigned long long magic = 0x6d676c6300000000ull + (unsigned long long)h->rank;
    ok &= cudaMemcpy(h->flags + 31, &magic, sizeof magic, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int b = 0; b < 2; ++b) {
        ok &= cudaIpcGetMemHandle(&mine.buf[b], h->buf[b]) == cudaSuccess;
        if (h->thermal) ok &= cudaIpcGetMemHandle(&mine.gbuf[b], h->gbuf[b]) == cudaSuccess;
    }
    ok &= cudaIpcGetMemHandle(&mine.flags, h->flags) == cudaSuccess;
    if (h->xstage) { ok &= cudaIpcGetMemHandle(&mine.xstage, h->xstage) == cudaSuccess; mine.has_xstage = 1; }
    (void)cudaGetLastError();
    for (int q = 0; q < 3; ++q) mine.ln[q] = h->d.ln[q];
    mine.thermal = h->thermal;
    char *dev_all = nullptr;
    MGLC_CUDA(cudaMalloc((void **)&dev_all, (size_t)(P + 1) * sizeof(IpcRecord)));
    MGLC_CUDA(cudaMemcpy(dev_all + (size_t)P * sizeof(IpcRecord), &mine, sizeof mine, cudaMemcpyHostToDevice));
    MGLC_NCCL(ncclAllGather(dev_all + (size_t)P * sizeof(IpcRecord), dev_all, sizeof(IpcRecord), ncclChar, h->comm->nccl, h->s));
    std::vector<IpcRecord> all(P);
    MGLC_CUDA(cudaMemcpyAsync(all.data(), dev_all, (size_t)P * sizeof(IpcRecord), cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    cudaFree(dev_all);
    h->ipc_opened = new std::vector<void *>();
    std::vector<PeerView> by_rank(P);
    std::vector<char> have(P, 0);
    PeerView view[19];
    memset(view, 0, sizeof view);
    auto open = [&](const cudaIpcMemHandle_t &hd, void **out) {
        if (cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); *out = nullptr; return 0; }
        h->ipc_opened->push_back(*out);
        return 1;
    };
    for (int d = 0; d < 19 && ok; ++d) {
        const int r = h->nbr_rank[d];
        if (d == 6 || r < 0) continue;
        if (!have[r]) {
            PeerView &v = by_rank[r];
            memset(&v, 0, sizeof v);
            for (int b = 0; b < 2 && ok; ++b) {
                ok &= open(all[r].buf[b], (void **)&v.buf[b]);
                if (h->thermal && ok) ok &= open(all[r].gbuf[b], (void **)&v.gbuf[b]);
            }
            if (ok) ok &= open(all[r].flags, (void **)&v.flags);
            if (ok && all[r].has_xstage) ok &= open(all[r].xstage, (void **)&v.xstage);
            if (ok) {           // the mapping must start at the neighbour's own pointer, not at some enclosing block
                unsigned long long seen = 0;
                ok &= cudaMemcpy(&seen, v.flags + 31, sizeof seen, cudaMemcpyDeviceToHost) == cudaSuccess &&
                      seen == 0x6d676c6300000000ull + (unsigned long long)r;
                (void)cudaGetLastError();
            }
            for (int q = 0; q < 3; ++q) v.ln[q] = all[r].ln[q];
            have[r] = 1;
        }
        view[d] = by_rank[r];
    }
    // collective verdict: everybody or nobody
    int *dev_ok = nullptr;
    MGLC_CUDA(cudaMalloc((void **)&dev_ok, sizeof(int)));
    MGLC_CUDA(cudaMemcpy(dev_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
    MGLC_NCCL(ncclAllReduce(dev_ok, dev_ok, 1, ncclInt, ncclMin, h->comm->nccl, h->s));
    MGLC_CUDA(cudaMemcpyAsync(&ok, dev_ok, sizeof ok, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    cudaFree(dev_ok);
    if (!ok) return MGLC_OK;                     // stay on the NCCL transport
    MGLC_TRY(install_peers(h, view));
    h->overlap = default_direct_mode(h);
    return MGLC_OK;
}

extern "C" int mglc_lbm_create(mglc_lbm **h, const mglc_lbm_desc *d, mglc_comm *comm_or_null) {
    if (d && d->dims[0] * d->dims[1] * d->dims[2] > 1 && !comm_or_null) {
        set_error("mglc_lbm_create: a %dx%dx%d decomposition needs a communicator (or use mglc_group_create)",
                  d->dims[0], d->dims[1], d->dims[2]);
        return MGLC_E_INVALID;
    }
    MGLC_TRY(create_impl(h, d, comm_or_null));
    if (comm_or_null && (*h)->has_neighbors) {
        const int rc = setup_direct_ipc(*h);        // collective over the communicator; failure to map = NCCL transport
        if (rc) { mglc_lbm_destroy(*h); *h = nullptr; return rc; }
    }
    return MGLC_OK;
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
// The neighbour barrier of the direct-halo path reports through a sticky device word (k_halo_wait).  Every entry point
// that hands results to the caller reads it once its stream has drained: a run whose neighbours fell out of step never
// comes back as MGLC_OK.
static int direct_status(mglc_lbm *h) {
    if (!h->direct || h->group) return MGLC_OK;
    int e = 0;
    MGLC_CUDA(cudaMemcpy(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e & 1) { set_error("direct halo path: a neighbour did not reach the barrier within %.0f s (MGLC_HALO_TIMEOUT_S; ranks out of step?); nothing was stored into a neighbour from that step on and this subdomain's fields are invalid", halo_timeout_seconds()); return MGLC_E_STATE; }
    if (e & 2) { set_error("direct halo path: a neighbour wrote the other ping-pong lattice (calls between mglc_lbm_step must be made by every rank)"); return MGLC_E_STATE; }
    return MGLC_OK;
}
extern "C" int mglc_lbm_sync(mglc_lbm *h) {
    MGLC_TRY(use(h));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s_comm));
    MGLC_CUDA(cudaGetLastError());
    return direct_status(h);
}

static int canonicalise(mglc_lbm *h);

// ---- host <-> device transfers in the reference layout ---------------------------------------------------
// Two staging halves: the PCIe copy of chunk c+1 (stream s_comm) runs beside the AoS<->SoA transpose of chunk c (stream s),
// so the transfer runs at the copy rate (the transposes move 64 MiB in ~30 us; the copy of the same chunk takes ~1.2 ms).
static int ensure_stage(mglc_lbm *h) {
    if (h->stage) return MGLC_OK;
    h->stage_doubles = 8LL << 20;   // 2 x 32 MiB
    for (int b = 0; b < 2; ++b) {
        MGLC_CUDA(cudaEventCreateWithFlags(&h->ev_stage_full[b], cudaEventDisableTiming));
        MGLC_CUDA(cudaEventCreateWithFlags(&h->ev_stage_free[b], cudaEventDisableTiming));
    }
    return dmalloc(h, &h->stage, h->stage_doubles);
}
// f (0:18,nx,ny,nz) or f_post (0:18,0:nx+1,...) on the host <-> SoA lattice on the device
static int transfer_lattice(mglc_lbm *h, double *host, double *dev, int with_halo, bool to_device, int nq = Q) {
    MGLC_TRY(ensure_stage(h));
    const int Q = nq;
    const int e = with_halo ? 2 : 0;
    const long long total = (long long)(h->g.nx + e) * (h->g.ny + e) * (h->g.nz + e);
    const long long half = h->stage_doubles / 2;
    const long long chunk = half / Q;
    // everything queued on s so far (canonicalise, earlier copies) comes before the first chunk on the copy stream
    MGLC_CUDA(cudaEventRecord(h->ev_stage_free[0], h->s));
    MGLC_CUDA(cudaEventRecord(h->ev_stage_free[1], h->s));
    int b = 0;
    for (long long c0 = 0; c0 < total; c0 += chunk, b ^= 1) {
        const long long nc = std::min(chunk, total - c0);
        double *st = h->stage + b * half;
        if (to_device) {
            MGLC_CUDA(cudaStreamWaitEvent(h->s_comm, h->ev_stage_free[b], 0));      // the transpose that read this half last
            MGLC_CUDA(cudaMemcpyAsync(st, host + c0 * Q, (size_t)nc * Q * sizeof(double), cudaMemcpyHostToDevice, h->s_comm));
            MGLC_CUDA(cudaEventRecord(h->ev_stage_full[b], h->s_comm));
            MGLC_CUDA(cudaStreamWaitEvent(h->s, h->ev_stage_full[b], 0));
            h->launches += launch_aos_to_soa(h->g, nq, st, dev, c0, nc, with_halo, h->s);
            MGLC_CUDA(cudaEventRecord(h->ev_stage_free[b], h->s));
        } else {
            MGLC_CUDA(cudaStreamWaitEvent(h->s, h->ev_stage_free[b], 0));           // the copy that drained this half last
            h->launches += launch_soa_to_aos(h->g, nq, dev, st, c0, nc, with_halo, h->s);
            MGLC_CUDA(cudaEventRecord(h->ev_stage_full[b], h->s));
            MGLC_CUDA(cudaStreamWaitEvent(h->s_comm, h->ev_stage_full[b], 0));
            MGLC_CUDA(cudaMemcpyAsync(host + c0 * Q, st, (size_t)nc * Q * sizeof(double), cudaMemcpyDeviceToHost, h->s_comm));
            MGLC_CUDA(cudaEventRecord(h->ev_stage_free[b], h->s_comm));
        }
    }
    MGLC_CUDA(cudaStreamSynchronize(h->s_comm));
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
    return direct_status(h);
}
extern "C" int mglc_lbm_download_f(mglc_lbm *h, double *f) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (!f) return MGLC_E_INVALID;
    MGLC_TRY(transfer_lattice(h, f, F_(h), 0, false));
    return direct_status(h);
}
extern "C" int mglc_lbm_download_fpost(mglc_lbm *h, double *f_post) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (!f_post) return MGLC_E_INVALID;
    MGLC_TRY(transfer_lattice(h, f_post, Fpost_(h), 1, false));
    return direct_status(h);
}

// ---- single-subdomain building blocks ---------------------------------------------------------------------
// before a subdomain is re-initialised: the neighbours' stores into my halos (direct path) or the exchange on s_comm
// (overlapped path) that a previous step() left in flight must have landed, or they would hit the fresh lattice
static int drain_halos(mglc_lbm *h) {
    if (h->direct_valid) MGLC_TRY(wait_direct(h));
    if (h->halo_inflight) {
        MGLC_CUDA(cudaStreamWaitEvent(h->s, h->ev_halo, 0));
        h->halo_inflight = 0;
    }
    return MGLC_OK;
}
static int do_initial(mglc_lbm *h) {
    MGLC_TRY(drain_halos(h));
    h->rotated = 0;
    h->lid_next = h->lid_last_in = h->rho + (long long)h->g.nx * h->g.ny * (h->g.nz - 1);
    if (h->thermal) {
        h->launches += launch_th_initial(h->g, h->tp, F_(h), G_(h), h->rho, h->u, h->v, h->w, h->T, h->s);
        MGLC_CUDA(cudaMemsetAsync(Fpost_(h), 0, (size_t)Q * h->g.sq * sizeof(double), h->s));      // f_post = 0, B3:624
        MGLC_CUDA(cudaMemsetAsync(Gpost_(h), 0, (size_t)QT * h->g.sq * sizeof(double), h->s));     // g_post = 0, B3:625
        if (h->Tp) MGLC_CUDA(cudaMemsetAsync(h->Tp, 0, (size_t)ncell(h) * sizeof(double), h->s));
    } else
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
    if (h->thermal) {
        h->launches += strict_(h) ? strict::launch_th_collision(h->g, h->tp, F_(h), h->rho, h->u, h->v, h->w, h->T, Fpost_(h), h->Fc[h->fc_cur], h->s)
                                  : fast::launch_th_collision(h->g, h->tp, F_(h), h->rho, h->u, h->v, h->w, h->T, Fpost_(h), h->Fc[h->fc_cur], h->s);
        return MGLC_OK;
    }
    h->launches += strict_(h) ? strict::launch_collision(h->g, h->p, F_(h), h->rho, h->u, h->v, h->w, Fpost_(h), h->s)
                              : fast::launch_collision(h->g, h->p, F_(h), h->rho, h->u, h->v, h->w, Fpost_(h), h->s);
    return MGLC_OK;
}
// which = MSG_F (f messages), MSG_G (g messages) or MSG_ALL: mark the others as skipped
static void select_msgs(mglc_lbm *h, int which) {
    for (int m = 0; m < h->nmsgs; ++m) h->msgs[m].skip = !(((h->msgs[m].dir >= 20) ? MSG_G : MSG_F) & which);
}
static int do_pack(mglc_lbm *h, cudaStream_t s) {
    MsgBatch mb{};
    for (int m = 0; m < h->nmsgs; ++m) {
        const Msg &M = h->msgs[m];
        if (!M.send_count || M.skip) continue;
        mb.dir[mb.n] = M.dir; mb.buf[mb.n] = M.sbuf; ++mb.n;
    }
    h->launches += launch_pack_all(h->g, mb, Fpost_(h), h->thermal ? Gpost_(h) : nullptr, false, s);
    return MGLC_OK;
}
static int do_unpack(mglc_lbm *h, cudaStream_t s) {
    MsgBatch mb{};
    for (int m = 0; m < h->nmsgs; ++m) {
        const Msg &M = h->msgs[m];
        if (!M.recv_count || M.skip) continue;
        mb.dir[mb.n] = M.dir; mb.buf[mb.n] = M.rbuf; ++mb.n;
    }
    h->launches += launch_pack_all(h->g, mb, Fpost_(h), h->thermal ? Gpost_(h) : nullptr, true, s);
    return MGLC_OK;
}
// message_passing_sendrecv() for a handle that owns a communicator (one process per GPU)
static int do_exchange_nccl(mglc_lbm *h, int which = MSG_ALL) {
    if (!h->has_neighbors) return MGLC_OK;
    if (!h->comm) { set_error("exchange: subdomain has neighbours but no communicator"); return MGLC_E_STATE; }
    select_msgs(h, which);
    MGLC_TRY(do_pack(h, h->s));
    MGLC_TRY(halo_nccl_sendrecv(h->msgs, h->nmsgs, h->comm, h->s));
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
    if (h->thermal) h->launches += launch_th_macro(h->g, F_(h), h->Fc[h->fc_cur], h->rho, h->u, h->v, h->w, h->s);
    else h->launches += launch_macro(h->g, F_(h), h->rho, h->u, h->v, h->w, h->s);
    return MGLC_OK;
}
static int need_thermal(mglc_lbm *h, const char *what) {
    if (!h->thermal) { set_error("%s: the handle is not a thermal (MGLC_D3Q19_D3Q7) lattice", what); return MGLC_E_STATE; }
    return MGLC_OK;
}
static int do_collisionT(mglc_lbm *h) {
    MGLC_TRY(need_thermal(h, "collisionT"));
    MGLC_TRY(canonicalise(h));
    h->launches += strict_(h) ? strict::launch_th_collisionT(h->g, h->tp, G_(h), h->u, h->v, h->w, h->T, Gpost_(h), h->s)
                              : fast::launch_th_collisionT(h->g, h->tp, G_(h), h->u, h->v, h->w, h->T, Gpost_(h), h->s);
    return MGLC_OK;
}
static int do_streamingT(mglc_lbm *h) {
    MGLC_TRY(need_thermal(h, "streamingT"));
    MGLC_TRY(canonicalise(h));
    h->launches += launch_streamingT(h->g, Gpost_(h), G_(h), h->s);
    return MGLC_OK;
}
static int do_bouncebackT(mglc_lbm *h) {
    MGLC_TRY(need_thermal(h, "bouncebackT"));
    MGLC_TRY(canonicalise(h));
    h->launches += launch_bouncebackT(h->g, h->tp, Gpost_(h), G_(h), h->s);
    return MGLC_OK;
}
static int do_macroT(mglc_lbm *h) {
    MGLC_TRY(need_thermal(h, "macroT"));
    MGLC_TRY(canonicalise(h));
    h->launches += launch_macroT(h->g, G_(h), h->T, h->s);
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
// one launch of the fused kernel over `box`, reading lattice `in` (with halos) and writing `out`; the lid plane /
// force buffers are the step's (see fused_begin)
struct FusedIO { const double *Fin, *Gin, *lid_in, *Fc_in; double *Fout, *Gout, *lid_out, *Fc_out; };
static int launch_fused_box(mglc_lbm *h, const FusedIO &io, const int box[6]) {
    if (h->thermal)
        h->launches += strict_(h) ? strict::launch_th_fused(h->g, h->tp, io.Fin, io.Fout, io.Gin, io.Gout, io.Fc_in, io.Fc_out, box, h->s)
                                  : fast::launch_th_fused(h->g, h->tp, io.Fin, io.Fout, io.Gin, io.Gout, io.Fc_in, io.Fc_out, box, h->s);
    else
        h->launches += strict_(h) ? strict::launch_fused(h->g, h->p, io.Fin, io.Fout, io.lid_in, io.lid_out, box, h->s)
                                  : fast::launch_fused(h->g, h->p, io.Fin, io.Fout, io.lid_in, io.lid_out, box, h->s);
    return MGLC_OK;
}
// buffers of the coming fused step + per-step profiling event; fused_end() rotates them
static int fused_begin(mglc_lbm *h, FusedIO &io) {
    if (h->profiling) {
        if (!h->prof_ev) {
            h->prof_ev = new std::vector<cudaEvent_t>(2 * PROF_PAIRS);
            for (cudaEvent_t &e : *h->prof_ev) MGLC_CUDA(cudaEventCreate(&e));
        }
        if (h->prof_used == PROF_PAIRS) MGLC_TRY(prof_flush(h));
        MGLC_CUDA(cudaEventRecord((*h->prof_ev)[2 * h->prof_used], h->s));
    }
    io = FusedIO{};
    io.Fin = Fpost_(h); io.Fout = F_(h);
    if (h->thermal) {
        io.Gin = Gpost_(h); io.Gout = G_(h);
        io.Fc_in = h->Fc[h->fc_cur]; io.Fc_out = h->Fc[h->fc_cur ^ 1];
    } else {
        io.lid_in = h->lid_next;
        io.lid_out = (io.lid_in == h->rho_lid[0]) ? h->rho_lid[1] : h->rho_lid[0];
    }
    return MGLC_OK;
}
// roles swap: the lattice just written becomes f_post of the next step
static void fused_swap(mglc_lbm *h, const FusedIO &io) {
    if (h->thermal) h->fc_cur ^= 1;   // Fc[fc_cur] now belongs to the collision in flight, Fc[fc_cur^1] to the completed step
    else { h->lid_last_in = io.lid_in; h->lid_next = io.lid_out; }
    h->cur ^= 1;
}
static int fused_end(mglc_lbm *h) {
    if (h->profiling) {
        MGLC_CUDA(cudaEventRecord((*h->prof_ev)[2 * h->prof_used + 1], h->s));
        h->prof_used += 1;
    }
    return MGLC_OK;
}
// fused stream+macro+collide over the whole subdomain: reads f_post (incl. halo), writes the NEXT f_post into the
// buffer that held f, then the two buffers swap roles
static int do_fused(mglc_lbm *h) {
    const int box[6] = {1, h->g.nx, 1, h->g.ny, 1, h->g.nz};
    FusedIO io;
    MGLC_TRY(fused_begin(h, io));
    MGLC_TRY(launch_fused_box(h, io, box));
    fused_swap(h, io);
    return fused_end(h);
}
// Direct-halo step: one launch over the whole subdomain that also stores the outgoing populations of its boundary
// cells into the neighbours' halos over NVLink (PeerTable), then the signal half of the neighbour barrier.  Nothing
// is packed, sent or unpacked; the transfer rides along with the update, tile by tile.
static int do_fused_direct(mglc_lbm *h) {
    const int box[6] = {1, h->g.nx, 1, h->g.ny, 1, h->g.nz};
    FusedIO io;
    MGLC_TRY(fused_begin(h, io));
    const PeerTable *pt = h->pt_dev[h->cur];                 // io.Fout == buf[cur]: the neighbours write the same parity
    if (h->overlap == 3) {
        // push after: the plain fused kernel, then one small launch that copies the messages into the neighbours' halos
        MGLC_TRY(launch_fused_box(h, io, box));
        h->launches += launch_push_halos(h->g, pt, io.Fout, h->thermal ? io.Gout : nullptr, h->s);
        h->xs_pending = h->xstage != nullptr;                // my x neighbours do the same: their x faces arrive in my staging
    } else if (h->thermal)
        h->launches += strict_(h) ? strict::launch_th_fused(h->g, h->tp, io.Fin, io.Fout, io.Gin, io.Gout, io.Fc_in, io.Fc_out, box, h->s, pt)
                                  : fast::launch_th_fused(h->g, h->tp, io.Fin, io.Fout, io.Gin, io.Gout, io.Fc_in, io.Fc_out, box, h->s, pt);
    else
        h->launches += strict_(h) ? strict::launch_fused(h->g, h->p, io.Fin, io.Fout, io.lid_in, io.lid_out, box, h->s, pt)
                                  : fast::launch_fused(h->g, h->p, io.Fin, io.Fout, io.lid_in, io.lid_out, box, h->s, pt);
    const int parity = h->cur;                               // of the lattice just written (before the swap)
    fused_swap(h, io);
    MGLC_TRY(fused_end(h));
    h->epoch += 1;
    h->parity_sent = parity;
    if (h->group) MGLC_CUDA(cudaEventRecord(h->ev_done[h->epoch & 1], h->s));
    else h->launches += launch_halo_signal(h->sync, h->epoch * 2 + (unsigned long long)parity, h->s);
    h->direct_valid = 1;
    return MGLC_OK;
}
// the wait half: every neighbour has finished the launch of my current epoch (its stores into my halos are complete,
// and it no longer reads the halos my next launch will overwrite)
static mglc_lbm *group_member(mglc_group *g, int r);
static int wait_direct(mglc_lbm *h) {
    if (h->group) {
        for (int d = 0; d < 19; ++d) {
            if (d == 6 || h->nbr_rank[d] < 0) continue;
            MGLC_CUDA(cudaStreamWaitEvent(h->s, group_member(h->group, h->nbr_rank[d])->ev_done[h->epoch & 1], 0));
        }
    } else h->launches += launch_halo_wait(h->sync, h->epoch * 2 + (unsigned long long)h->parity_sent, h->d_err, h->s);
    h->direct_valid = 0;
    if (h->xs_pending) {
        // halo push: the x-face messages of the lattice written last sit in my staging; scatter them into its halo columns
        // (message 0 came from the -x neighbour, message 1 from the +x neighbour)
        const int b = h->parity_sent;
        const long long per = (long long)h->g.ny * h->g.nz;
        MsgBatch mb{};
        for (int d = 0; d < 2; ++d) {
            if (h->nbr_rank[d ^ 1] < 0) continue;
            double *st = h->xstage + (b * 2 + d) * h->xs_count;
            mb.dir[mb.n] = d; mb.buf[mb.n] = st; ++mb.n;
            if (h->thermal) { mb.dir[mb.n] = 20 + d; mb.buf[mb.n] = st + 5 * per; ++mb.n; }
        }
        h->launches += launch_pack_all(h->g, mb, h->buf[b], h->thermal ? h->gbuf[b] : nullptr, true, h->s);
        h->xs_pending = 0;
    }
    return MGLC_OK;
}
// Overlapped step (the schedule of L3nb collision_with_message_exchange, :1108-1230): the cells next to a
// NEIGHBOUR face first (thin slabs: one plane in y and z, 32 cells in x so a warp still stores full lines),
// then their populations are packed, exchanged and unpacked on s_comm while the interior runs on s.
static int shell_x() {
    static int v = 0;
    if (!v) { v = 32; if (const char *e = getenv("MGLC_SHELL_X")) v = std::max(1, std::min(128, atoi(e))); }
    return v;
}
static int do_fused_overlapped(mglc_lbm *h) {
    const Geom &g = h->g;
    const int n[3] = {g.nx, g.ny, g.nz};
    const int want[3] = {shell_x(), 1, 1};
    int lo[3], hi[3];                 // interior box per axis
    for (int d = 0; d < 3; ++d) {
        const bool nb_plus = !g.wall[2 * d], nb_minus = !g.wall[2 * d + 1];
        int tm = nb_minus ? std::min(want[d], n[d]) : 0, tp = nb_plus ? std::min(want[d], n[d]) : 0;
        if (tm + tp > n[d]) { tm = n[d]; tp = 0; }
        lo[d] = 1 + tm; hi[d] = n[d] - tp;
    }
    FusedIO io;
    MGLC_TRY(fused_begin(h, io));
    // disjoint shell slabs: z slabs span everything, y slabs the interior k range, x slabs the interior j,k range
    const int zs[2][2] = {{1, lo[2] - 1}, {hi[2] + 1, n[2]}}, ys[2][2] = {{1, lo[1] - 1}, {hi[1] + 1, n[1]}},
              xs[2][2] = {{1, lo[0] - 1}, {hi[0] + 1, n[0]}};
    for (int q = 0; q < 2; ++q) { const int box[6] = {1, n[0], 1, n[1], zs[q][0], zs[q][1]}; MGLC_TRY(launch_fused_box(h, io, box)); }
    for (int q = 0; q < 2; ++q) { const int box[6] = {1, n[0], ys[q][0], ys[q][1], lo[2], hi[2]}; MGLC_TRY(launch_fused_box(h, io, box)); }
    for (int q = 0; q < 2; ++q) { const int box[6] = {xs[q][0], xs[q][1], lo[1], hi[1], lo[2], hi[2]}; MGLC_TRY(launch_fused_box(h, io, box)); }
    MGLC_CUDA(cudaEventRecord(h->ev_shell, h->s));
    fused_swap(h, io);                // from here on Fpost_(h) is the lattice being written: pack/unpack address it
    MGLC_CUDA(cudaStreamWaitEvent(h->s_comm, h->ev_shell, 0));
    select_msgs(h, MSG_ALL);
    MGLC_TRY(do_pack(h, h->s_comm));
    MGLC_TRY(halo_nccl_sendrecv(h->msgs, h->nmsgs, h->comm, h->s_comm));
    MGLC_TRY(do_unpack(h, h->s_comm));
    MGLC_CUDA(cudaEventRecord(h->ev_halo, h->s_comm));
    h->halo_inflight = 1;
    const int box[6] = {lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]};
    MGLC_TRY(launch_fused_box(h, io, box));
    return fused_end(h);
}
// make the halos of f_post (and g_post) of the coming step valid on stream s
static int halos_for_next_step(mglc_lbm *h) {
    if (h->halo_inflight) {
        MGLC_CUDA(cudaStreamWaitEvent(h->s, h->ev_halo, 0));
        h->halo_inflight = 0;
        return MGLC_OK;
    }
    return MGLC_OK;
}
static int do_stream_macro(mglc_lbm *h) {
    if (h->thermal) {
        const double *Fc_done = h->Fc[h->fc_cur ^ 1];        // the force the completed step's collision computed
        h->launches += strict_(h) ? strict::launch_th_stream_macro(h->g, h->tp, Fpost_(h), F_(h), Gpost_(h), G_(h), Fc_done, h->rho, h->u, h->v, h->w, h->T, h->s)
                                  : fast::launch_th_stream_macro(h->g, h->tp, Fpost_(h), F_(h), Gpost_(h), G_(h), Fc_done, h->rho, h->u, h->v, h->w, h->T, h->s);
        h->fc_cur ^= 1;
        return MGLC_OK;
    }
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
    if (h->halo_inflight) {           // the exchange of the discarded lattice must be off the wire before it is reused
        MGLC_CUDA(cudaStreamWaitEvent(h->s, h->ev_halo, 0));
        h->halo_inflight = 0;
    }
    if (h->direct_valid) MGLC_TRY(wait_direct(h));
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
    if (h->thermal) {
        MGLC_TRY(dmalloc(h, &h->Tp, n));
        MGLC_CUDA(cudaMemsetAsync(h->Tp, 0, (size_t)n * sizeof(double), h->s));      // Tp = 0, B3:619
    }
    return MGLC_OK;
}
static int do_check_partial(mglc_lbm *h) {
    MGLC_TRY(canonicalise(h));
    MGLC_TRY(ensure_prev(h));
    if (h->thermal) h->launches += launch_th_check(h->g, h->u, h->v, h->w, h->T, h->up, h->vp, h->wp, h->Tp, h->scratch, h->s);
    else h->launches += launch_check(h->g, h->u, h->v, h->w, h->up, h->vp, h->wp, h->scratch, h->s);
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
    return do_exchange_nccl(h, MSG_F);
}
extern "C" int mglc_exchange_g(mglc_lbm *h) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_exchange_g"));
    MGLC_TRY(need_thermal(h, "mglc_exchange_g"));
    MGLC_TRY(canonicalise(h));
    return do_exchange_nccl(h, MSG_G);
}
extern "C" int mglc_collisionT(mglc_lbm *h) { MGLC_TRY(use(h)); return do_collisionT(h); }
extern "C" int mglc_streamingT(mglc_lbm *h) { MGLC_TRY(use(h)); return do_streamingT(h); }
extern "C" int mglc_bouncebackT(mglc_lbm *h) { MGLC_TRY(use(h)); return do_bouncebackT(h); }
extern "C" int mglc_macroT(mglc_lbm *h) { MGLC_TRY(use(h)); return do_macroT(h); }
extern "C" int mglc_streaming(mglc_lbm *h) { MGLC_TRY(use(h)); return do_streaming(h); }
extern "C" int mglc_bounceback(mglc_lbm *h) { MGLC_TRY(use(h)); return do_bounceback(h); }
extern "C" int mglc_macro(mglc_lbm *h) { MGLC_TRY(use(h)); return do_macro(h); }

static int check_impl(mglc_lbm *h, double *errorU, double *errorT) {
    MGLC_TRY(do_check_partial(h));
    if (h->comm && h->nranks > 1)   // MPI_Allreduce(SUM) x2 (L3/check.f90:27-28) or x4 (B3:1268-1271)
        MGLC_NCCL(ncclAllReduce(h->scratch, h->scratch, 4, ncclDouble, ncclSum, h->comm->nccl, h->s));
    double e[4];
    MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    *errorU = sqrt(e[0]) / sqrt(e[1]);
    if (errorT) *errorT = e[2] / e[3];
    return direct_status(h);
}
extern "C" int mglc_check(mglc_lbm *h, double *errorU) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_check"));
    if (!errorU) return MGLC_E_INVALID;
    return check_impl(h, errorU, nullptr);
}
extern "C" int mglc_check_thermal(mglc_lbm *h, double *errorU, double *errorT) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_check_thermal"));
    MGLC_TRY(need_thermal(h, "mglc_check_thermal"));
    if (!errorU || !errorT) return MGLC_E_INVALID;
    return check_impl(h, errorU, errorT);
}
// calNuRe(): B3/mpi_blocked/RaNu.F90:13-47.  The reference's subroutine sums over one rank's block (its call is commented
// out in the MPI driver, B3:264); here the two sums are all-reduced and the averages taken over the global box, which is
// what the sequential driver computes (B3/seq/bouyancy3d.F90, calNuRe).
static void nure_from_sums(const mglc_lbm *h, double prandtl, const double sums[2], double *Nu, double *Re) {
    const double viscosity = (h->d.tau - 0.5) / 3.0, diffusivity = viscosity / prandtl;      // module commondata, B3:38-39
    const double ncells = (double)((long long)h->d.gn[0] * h->d.gn[1] * h->d.gn[2]), nz = (double)h->d.gn[2];
    *Nu = sums[0] / ncells * nz / diffusivity + 1.0;
    *Re = sqrt(sums[1] / ncells) * nz / viscosity;
}
static int do_nure_partial(mglc_lbm *h) {
    MGLC_TRY(need_thermal(h, "mglc_calNuRe"));
    MGLC_TRY(canonicalise(h));
    h->launches += launch_nure(h->g, h->u, h->v, h->w, h->T, h->scratch, h->s);
    return MGLC_OK;
}
extern "C" int mglc_calNuRe(mglc_lbm *h, double prandtl, double *NuVolAvg, double *ReVolAvg) {
    MGLC_TRY(use(h));
    MGLC_TRY(not_in_group(h, "mglc_calNuRe"));
    if (!NuVolAvg || !ReVolAvg || !(prandtl > 0.0)) { set_error("mglc_calNuRe: bad arguments"); return MGLC_E_INVALID; }
    MGLC_TRY(do_nure_partial(h));
    if (h->comm && h->nranks > 1) MGLC_NCCL(ncclAllReduce(h->scratch, h->scratch, 2, ncclDouble, ncclSum, h->comm->nccl, h->s));
    double e[2];
    MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    nure_from_sums(h, prandtl, e, NuVolAvg, ReVolAvg);
    return direct_status(h);
}
// one line of a macroscopic field along `axis` through the global 1-based indices (g1, g2) of the other two axes (ascending
// axis order), e.g. u(nxHalf, nyHalf, :) of getVelocity(), L3/output.f90:334-344, without a full-field download.
// field: 0 rho, 1 u, 2 v, 3 w, 4 T.  A subdomain the line does not cross returns *count = 0.
extern "C" int mglc_lbm_download_line(mglc_lbm *h, int field, int axis, int g1, int g2, double *out, int *first, int *count) {
    MGLC_TRY(use(h));
    if (field < 0 || field > 4 || axis < 0 || axis > 2 || !out || !first || !count) { set_error("mglc_lbm_download_line: bad arguments"); return MGLC_E_INVALID; }
    if (field == 4) MGLC_TRY(need_thermal(h, "mglc_lbm_download_line"));
    MGLC_TRY(canonicalise(h));
    const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
    const int l1 = g1 - h->d.start[a1], l2 = g2 - h->d.start[a2];      // 1-based local
    *first = h->d.start[axis] + 1;
    if (l1 < 1 || l1 > h->d.ln[a1] || l2 < 1 || l2 > h->d.ln[a2]) { *count = 0; return MGLC_OK; }
    const double *src = field == 0 ? h->rho : field == 1 ? h->u : field == 2 ? h->v : field == 3 ? h->w : h->T;
    const long long stride[3] = {1, h->g.nx, (long long)h->g.nx * h->g.ny};
    const double *p0 = src + (l1 - 1) * stride[a1] + (l2 - 1) * stride[a2];
    const int n = h->d.ln[axis];
    MGLC_CUDA(cudaMemcpy2DAsync(out, sizeof(double), p0, (size_t)stride[axis] * sizeof(double), sizeof(double), (size_t)n,
                                cudaMemcpyDeviceToHost, h->s));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    *count = n;
    return direct_status(h);
}
extern "C" int mglc_lbm_upload_thermal(mglc_lbm *h, const double *g, const double *T, const double *Fx, const double *Fy,
                                       const double *Fz) {
    MGLC_TRY(use(h));
    MGLC_TRY(need_thermal(h, "mglc_lbm_upload_thermal"));
    MGLC_TRY(canonicalise(h));
    if (g) MGLC_TRY(transfer_lattice(h, const_cast<double *>(g), G_(h), 0, true, QT));
    MGLC_TRY(copy_field(h, const_cast<double *>(T), h->T, true));
    double *Fc = h->Fc[h->fc_cur];
    MGLC_TRY(copy_field(h, const_cast<double *>(Fx), Fc, true));
    MGLC_TRY(copy_field(h, const_cast<double *>(Fy), Fc + ncell(h), true));
    MGLC_TRY(copy_field(h, const_cast<double *>(Fz), Fc + 2 * ncell(h), true));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return MGLC_OK;
}
extern "C" int mglc_lbm_download_thermal(mglc_lbm *h, double *g, double *T, double *Fx, double *Fy, double *Fz) {
    MGLC_TRY(use(h));
    MGLC_TRY(need_thermal(h, "mglc_lbm_download_thermal"));
    MGLC_TRY(canonicalise(h));
    if (g) MGLC_TRY(transfer_lattice(h, g, G_(h), 0, false, QT));
    MGLC_TRY(copy_field(h, T, h->T, false));
    double *Fc = h->Fc[h->fc_cur];
    MGLC_TRY(copy_field(h, Fx, Fc, false));
    MGLC_TRY(copy_field(h, Fy, Fc + ncell(h), false));
    MGLC_TRY(copy_field(h, Fz, Fc + 2 * ncell(h), false));
    MGLC_CUDA(cudaStreamSynchronize(h->s));
    return direct_status(h);
}
extern "C" int mglc_lbm_upload_gpost(mglc_lbm *h, const double *g_post) {
    MGLC_TRY(use(h));
    MGLC_TRY(need_thermal(h, "mglc_lbm_upload_gpost"));
    MGLC_TRY(canonicalise(h));
    if (!g_post) return MGLC_E_INVALID;
    return transfer_lattice(h, const_cast<double *>(g_post), Gpost_(h), 1, true, QT);
}
extern "C" int mglc_lbm_download_gpost(mglc_lbm *h, double *g_post) {
    MGLC_TRY(use(h));
    MGLC_TRY(need_thermal(h, "mglc_lbm_download_gpost"));
    MGLC_TRY(canonicalise(h));
    if (!g_post) return MGLC_E_INVALID;
    MGLC_TRY(transfer_lattice(h, g_post, Gpost_(h), 1, false, QT));
    return direct_status(h);
}

// nsteps iterations of: collision, exchange, streaming, bounceback, macro  (L3/main.f90:85-97).
// Rotated by half a step: [collision once] then nsteps x (exchange -> fused pull+walls+macro+collide).
// The handle stays rotated afterwards, so back-to-back calls pay neither prologue nor epilogue; any
// entry point that needs the reference's state calls canonicalise() first.
static int step_impl(mglc_lbm *h, int nsteps) {
    if (nsteps < 0) { set_error("mglc_lbm_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    if (nsteps == 0) return MGLC_OK;
    if (!h->rotated) {                               // the first step's collision() (and collisionT())
        MGLC_TRY(do_collision(h));
        if (h->thermal) MGLC_TRY(do_collisionT(h));
    }
    const bool direct = (h->overlap == 2 || h->overlap == 3) && h->direct && h->has_neighbors && h->comm;
    const bool overlapped = h->overlap == 1 && h->has_neighbors && h->comm;
    for (int it = 0; it < nsteps; ++it) {
        if (h->direct_valid) MGLC_TRY(wait_direct(h));            // the neighbours' launches stored the halos of step it
        else if (h->halo_inflight) MGLC_TRY(halos_for_next_step(h));   // exchange of step it already ran beside the last interior
        else MGLC_TRY(do_exchange_nccl(h));                       // exchange of step it
        if (direct) MGLC_TRY(do_fused_direct(h));
        else if (overlapped) MGLC_TRY(do_fused_overlapped(h));
        else MGLC_TRY(do_fused(h));                  // streaming+bounceback+macro of step it, collision of it+1
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
    return direct_status(h);
}
// 1 (default) = overlap the halo exchange with the interior update when the handle has neighbours and a
// communicator; 0 = exchange, then update (what the blocking reference driver does, L3/main.f90:89-93)
extern "C" int mglc_lbm_set_overlap(mglc_lbm *h, int on) {
    MGLC_TRY(use(h));
    MGLC_TRY(canonicalise(h));
    if (on < 0 || on > 3) { set_error("mglc_lbm_set_overlap: mode %d (0 blocking, 1 overlapped exchange, 2 direct halo stores, 3 halo push after the update)", on); return MGLC_E_INVALID; }
    if (on >= 2 && !h->direct && h->has_neighbors) { set_error("mglc_lbm_set_overlap: the neighbours' lattices are not mapped on this GPU"); return MGLC_E_STATE; }
    h->overlap = on;
    return MGLC_OK;
}
extern "C" int mglc_lbm_get_overlap(mglc_lbm *h, int *mode) {
    if (!h || !mode) return MGLC_E_INVALID;
    *mode = h->overlap;
    return MGLC_OK;
}
extern "C" int mglc_lbm_direct_halo(mglc_lbm *h, int *available) {
    if (!h || !available) return MGLC_E_INVALID;
    *available = h->direct;
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
    // quiesce every member before the first one is freed: neighbours store into each other's halos
    for (mglc_lbm *h : g->r) { cudaSetDevice(h->d.device); cudaDeviceSynchronize(); h->direct_valid = 0; }
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
    for (mglc_lbm *h : g->r) g->ports.push_back(Port{h->d.device, h->s, h->ev_packed, h->ev_copied, h->msgs, h->nmsgs});
    // direct halo stores inside one process: the neighbours' lattices are plain device pointers (same device, or peer
    // access enabled above); ordering goes through events instead of flag words
    bool reachable = nranks > 1 && !getenv("MGLC_NO_DIRECT");
    for (mglc_lbm *a : g->r)
        for (mglc_lbm *b : g->r)
            if (reachable && a->d.device != b->d.device) { int can = 0; cudaDeviceCanAccessPeer(&can, a->d.device, b->d.device); reachable = can != 0; }
    if (reachable)
        for (mglc_lbm *h : g->r) {
            PeerView view[19];
            memset(view, 0, sizeof view);
            for (int dd = 0; dd < 19; ++dd) {
                if (dd == 6 || h->nbr_rank[dd] < 0) continue;
                mglc_lbm *n = g->r[h->nbr_rank[dd]];
                for (int b = 0; b < 2; ++b) { view[dd].buf[b] = n->buf[b]; view[dd].gbuf[b] = n->gbuf[b]; }
                view[dd].flags = n->flags;
                view[dd].xstage = n->xstage;
                for (int q = 0; q < 3; ++q) view[dd].ln[q] = n->d.ln[q];
            }
            int rc = use(h);
            if (!rc) rc = install_peers(h, view);
            if (rc) { mglc_group_destroy(g); return rc; }
            h->overlap = default_direct_mode(h);
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

static mglc_lbm *group_member(mglc_group *g, int r) { return g->r[r]; }
#define FOR_RANKS(g, h) for (mglc_lbm * h : (g)->r)
#define GROUP_EACH(g, fn)                       \
    do {                                        \
        if (!(g)) return MGLC_E_INVALID;        \
        FOR_RANKS(g, h_) { MGLC_TRY(use(h_)); MGLC_TRY(fn(h_)); } \
        return MGLC_OK;                         \
    } while (0)

// message_passing_sendrecv() across the group: pack on every sender, receiver-driven device-to-device
// copies (ordered by events), unpack on every receiver
static int group_exchange(mglc_group *g, int which = MSG_ALL) {
    FOR_RANKS(g, h) select_msgs(h, which);
    return halo_local_exchange(
        g->ports, [&](int r, cudaStream_t s) { return do_pack(g->r[r], s); },
        [&](int r, cudaStream_t s) { return do_unpack(g->r[r], s); });
}

extern "C" int mglc_group_initial(mglc_group *g) {
    if (!g) return MGLC_E_INVALID;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(drain_halos(h)); }      // every member, before any member starts over
    GROUP_EACH(g, do_initial);
}
extern "C" int mglc_group_collision(mglc_group *g) { GROUP_EACH(g, do_collision); }
extern "C" int mglc_group_exchange(mglc_group *g) {
    if (!g) return MGLC_E_INVALID;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(canonicalise(h)); }
    return group_exchange(g, MSG_F);
}
extern "C" int mglc_group_exchange_g(mglc_group *g) {
    if (!g) return MGLC_E_INVALID;
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(need_thermal(h, "mglc_group_exchange_g")); MGLC_TRY(canonicalise(h)); }
    return group_exchange(g, MSG_G);
}
extern "C" int mglc_group_collisionT(mglc_group *g) { GROUP_EACH(g, do_collisionT); }
extern "C" int mglc_group_streamingT(mglc_group *g) { GROUP_EACH(g, do_streamingT); }
extern "C" int mglc_group_bouncebackT(mglc_group *g) { GROUP_EACH(g, do_bouncebackT); }
extern "C" int mglc_group_macroT(mglc_group *g) { GROUP_EACH(g, do_macroT); }
extern "C" int mglc_group_streaming(mglc_group *g) { GROUP_EACH(g, do_streaming); }
extern "C" int mglc_group_bounceback(mglc_group *g) { GROUP_EACH(g, do_bounceback); }
extern "C" int mglc_group_macro(mglc_group *g) { GROUP_EACH(g, do_macro); }

static int group_check_impl(mglc_group *g, double *errorU, double *errorT) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(do_check_partial(h)); }
    FOR_RANKS(g, h) {                       // Allreduce(SUM) modelled as a rank-ordered host sum
        MGLC_TRY(use(h));
        double e[4];
        MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
        MGLC_CUDA(cudaStreamSynchronize(h->s));
        for (int q = 0; q < 4; ++q) t[q] += e[q];
    }
    *errorU = sqrt(t[0]) / sqrt(t[1]);
    if (errorT) *errorT = t[2] / t[3];
    return MGLC_OK;
}
extern "C" int mglc_group_check(mglc_group *g, double *errorU) {
    if (!g || !errorU) return MGLC_E_INVALID;
    return group_check_impl(g, errorU, nullptr);
}
extern "C" int mglc_group_check_thermal(mglc_group *g, double *errorU, double *errorT) {
    if (!g || !errorU || !errorT) return MGLC_E_INVALID;
    FOR_RANKS(g, h) MGLC_TRY(need_thermal(h, "mglc_group_check_thermal"));
    return group_check_impl(g, errorU, errorT);
}

extern "C" int mglc_group_calNuRe(mglc_group *g, double prandtl, double *NuVolAvg, double *ReVolAvg) {
    if (!g || !NuVolAvg || !ReVolAvg || !(prandtl > 0.0)) return MGLC_E_INVALID;
    double t[2] = {0.0, 0.0};
    FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(do_nure_partial(h)); }
    FOR_RANKS(g, h) {                       // Allreduce(SUM) modelled as a rank-ordered host sum
        MGLC_TRY(use(h));
        double e[2];
        MGLC_CUDA(cudaMemcpyAsync(e, h->scratch, sizeof e, cudaMemcpyDeviceToHost, h->s));
        MGLC_CUDA(cudaStreamSynchronize(h->s));
        t[0] += e[0]; t[1] += e[1];
    }
    nure_from_sums(g->r[0], prandtl, t, NuVolAvg, ReVolAvg);
    return MGLC_OK;
}

static int group_step_impl(mglc_group *g, int nsteps) {
    if (nsteps < 0) return MGLC_E_INVALID;
    if (nsteps == 0) return MGLC_OK;
    // a member that was brought back to the reference's state on its own (per-subdomain download, ...) has flipped its
    // ping-pong parity relative to the others; direct stores need every member on the same parity
    bool mixed = false;
    FOR_RANKS(g, h) mixed = mixed || (h->rotated != g->r[0]->rotated);
    if (mixed) FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(canonicalise(h)); }
    FOR_RANKS(g, h) {
        MGLC_TRY(use(h));
        if (!h->rotated) { MGLC_TRY(do_collision(h)); if (h->thermal) MGLC_TRY(do_collisionT(h)); }
    }
    const bool direct = g->r[0]->direct && g->r[0]->overlap >= 2 && g->r.size() > 1;
    for (int it = 0; it < nsteps; ++it) {
        bool all_valid = true;
        FOR_RANKS(g, h) all_valid = all_valid && h->direct_valid;
        FOR_RANKS(g, h) if (h->direct_valid) { MGLC_TRY(use(h)); MGLC_TRY(wait_direct(h)); }
        if (!all_valid) MGLC_TRY(group_exchange(g));
        FOR_RANKS(g, h) { MGLC_TRY(use(h)); MGLC_TRY(direct ? do_fused_direct(h) : do_fused(h)); }
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
