// thermal2d.cu -- the reference's 2-D double-distribution thermal driver (D2Q9 MRT flow + Boussinesq force, D2Q5 MRT temperature;
// SURVEY 8f row 4) behind the mglc_t2d_* entry points of mglc.h:
//   B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/*.F90    (Fortran + MPI, 2-D Cartesian blocks, 201 x 201, Ra = 1e7)
// This file holds the strict build of the collision / fused kernels (-fmad=false), the copy-type subroutines (streaming,
// bounceback, streamingT, bouncebackT, macro, macroT, initial, check, the calNuRe sums, halo pack/unpack, layout transposes)
// and the host side; thermal2d_fast.cu is the throughput build of the same kernel source.
#define MGLC_NS strict
#define MGLC_STRICT 1
#include "thermal2d_kernels.inl"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "halo.cuh"

using namespace mglc;

namespace {

#include "thermal2d_exact.inl"

constexpr int T2_NMSG = 12;

struct T2Sub {
    int n[2], coords[2], start[2];
    int nbr[4];          // right(+x), left(-x), top(+y), bottom(-y); -1 = MPI_PROC_NULL     main.F90:41-42
    int cnr[4];          // the neighbours populations 5..8 travel to                        MPI_Cart_find_corners, main.F90:228-240
    int device;
    Geom2 g;
    T2Params p;          // the handle's parameters + this subdomain's place in the global lattice (start, total)
    double *F, *G;       // f, g        (pre-collision)
    double *P[2], *Q[2]; // f_post = P[cur], g_post = Q[cur]; the rotated loop ping-pongs between the two
    int cur;
    double *rho, *u, *v, *T, *up, *vp, *Tp, *Fx, *Fy;
    double *stage;       // reference-layout staging for upload / download
    double *scratch;     // reduction partial sums
    cudaStream_t s;
    cudaEvent_t ev_packed, ev_copied, ev_t0, ev_t1;
    Msg msgs[T2_NMSG];
    long long launches;
    cudaGraphExec_t gexec[2];      // T2_GRAPH_STEPS fused launches starting from cur = 0 / 1 (small single-subdomain lattices)
};

// A lattice that fits the L2 (the shipped 201 x 201 and 513 x 257 cases) is bounded by kernel-launch latency, not by HBM:
// such runs replay the fused launches as CUDA graphs of T2_GRAPH_STEPS kernels (an even count: the ping-pong index returns).
constexpr int T2_GRAPH_STEPS = 64;
// largest lattice (cells) that is replayed from graphs; MGLC_2D_GRAPH_CELLS overrides it (0 = never), for A/B measurements
static long long t2_graph_max_cells() {
    static long long v = -1;
    if (v < 0) { v = 1LL << 20; if (const char *e = getenv("MGLC_2D_GRAPH_CELLS")) v = std::max(0LL, atoll(e)); }
    return v;
}

}  // namespace

struct mglc_t2d {
    mglc_t2d_desc d;
    int dims[2], nranks;
    double lengthUnit, tauf, viscosity, diffusivity;
    T2Params p;
    std::vector<T2Sub *> subs;      // the subdomains this process owns (all of them, or exactly one)
    std::vector<Port> ports;
    mglc_comm *comm;
};

extern "C" int mglc_t2d_desc_init(mglc_t2d_desc *d) {
    if (!d) { set_error("mglc_t2d_desc_init: null descriptor"); return MGLC_E_INVALID; }
    memset(d, 0, sizeof *d);
    d->total_nx = d->total_ny = 201;                                            // module.F90:26
    d->arith = MGLC_ARITH_FAST;
    d->bcT[0] = MGLC_BCT_CONST_COLD; d->bcT[1] = MGLC_BCT_CONST_HOT;            // macros.F90:24-27: SideHeatedCell
    d->bcT[2] = d->bcT[3] = MGLC_BCT_ADIABATIC;
    d->Rayleigh = 1e7; d->Prandtl = 0.71; d->Mach = 0.1;                        // module.F90:31-33
    d->Thot = 1.0; d->Tcold = 0.0; d->Tref = 0.0; d->rho0 = 1.0;                // module.F90:67-68
    d->variant = MGLC_T2D_MPI; d->lengthUnit = 0.0;                             // lengthUnit = dble(total_ny), module.F90:29
    return MGLC_OK;
}
// the OpenACC program as shipped, seq/bouyancy2d_acc.F90 (the reference's only GPU code)
extern "C" int mglc_t2d_desc_init_acc(mglc_t2d_desc *d) {
    MGLC_TRY(mglc_t2d_desc_init(d));
    d->total_nx = 513; d->total_ny = 257;                                       // acc:55
    d->bcT[0] = d->bcT[1] = MGLC_BCT_PERIODIC;                                  // acc:15,22: VerticalWallsPeriodicalU / ...T
    d->bcT[2] = MGLC_BCT_CONST_COLD; d->bcT[3] = MGLC_BCT_CONST_HOT;            // acc:19-20: RayleighBenardCell, HorizontalWallsConstT
    d->Rayleigh = 1e5;                                                          // acc:59
    d->variant = MGLC_T2D_ACC; d->lengthUnit = 513.0;                           // acc:57: lengthUnit = dble(nx)
    return MGLC_OK;
}

// seq/R_B_2d.F90 as shipped (also the macro set of seq/bouyancy2d_omp.F90): Rayleigh-Benard plates, Pr = 5.3, every wall moving
// along itself at U0 = shearReynolds*viscosity/dble(ny) with shearReynolds = 100 (:61,79,118-120), its own corner cells in bouncebackT()
extern "C" int mglc_t2d_desc_init_sheared_rb(mglc_t2d_desc *d) {
    MGLC_TRY(mglc_t2d_desc_init(d));
    d->Prandtl = 5.3;                                                           // RB2:61
    d->bcT[0] = d->bcT[1] = MGLC_BCT_ADIABATIC;                                 // RB2:22-24: RayleighBenardCell
    d->bcT[2] = MGLC_BCT_CONST_COLD; d->bcT[3] = MGLC_BCT_CONST_HOT;
    const double lengthUnit = (double)d->total_ny;                              // RB2:58
    const double tauf = 0.5 + d->Mach * lengthUnit * sqrt(3.0 * d->Prandtl / d->Rayleigh), viscosity = (tauf - 0.5) / 3.0;   // :96-97
    const double U0 = 100.0 * viscosity / (double)d->total_ny;                  // :79,118
    const double w[8] = {U0, -U0, -U0, U0, U0, U0, U0, U0};                     // :119-120
    for (int q = 0; q < 8; ++q) d->Uwall[q] = w[q];
    d->cornersT = 1;                                                            // :1086-1106
    return MGLC_OK;
}

static int t2_use(T2Sub *S) { MGLC_CUDA(cudaSetDevice(S->device)); return MGLC_OK; }

static void t2_free_sub(T2Sub *S) {
    if (!S) return;
    cudaSetDevice(S->device);
    if (S->s) cudaStreamSynchronize(S->s);
    double *bufs[] = {S->F, S->G, S->P[0], S->P[1], S->Q[0], S->Q[1], S->rho, S->u, S->v, S->T, S->up, S->vp, S->Tp, S->Fx, S->Fy,
                      S->stage, S->scratch};
    for (double *p : bufs) cudaFree(p);
    for (Msg &M : S->msgs) { cudaFree(M.sbuf); cudaFree(M.rbuf); }
    cudaEvent_t evs[] = {S->ev_packed, S->ev_copied, S->ev_t0, S->ev_t1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (cudaGraphExec_t e : S->gexec) if (e) cudaGraphExecDestroy(e);
    if (S->s) cudaStreamDestroy(S->s);
    (void)cudaGetLastError();
    delete S;
}

extern "C" int mglc_t2d_destroy(mglc_t2d *h) {
    if (!h) return MGLC_OK;
    for (T2Sub *S : h->subs) t2_free_sub(S);
    delete h;
    return MGLC_OK;
}

static int t2_cart_rank(const int dims[2], int c0, int c1) {
    if (c0 < 0 || c0 >= dims[0] || c1 < 0 || c1 >= dims[1]) return -1;
    return c0 * dims[1] + c1;
}
static void t2_msg_dims(const T2Sub *S, int dir, int &n1, int &npop) {
    const int d = dir >= 8 ? dir - 8 : dir;
    if (d < 4) { n1 = (d >> 1) == 0 ? S->n[1] : S->n[0]; npop = dir >= 8 ? 1 : 3; }
    else { n1 = 1; npop = 1; }
}

// the 12 messages of message_passing_f() + message_passing_g() (message_exchange.F90:1-118) for the block at (c0, c1) with interior
// size n: pure host logic, also reachable without a GPU through mglc_t2d_msg_table (the CPU suite compares it with mglc_halo_plan_2d)
static void t2_build_msgs(const int dims[2], const int n[2], int c0, int c1, Msg msgs[T2_NMSG]) {
    for (int dir = 0; dir < T2_NMSG; ++dir) {
        Msg &M = msgs[dir];
        const int d = dir >= 8 ? dir - 8 : dir;
        const int n1 = d < 4 ? ((d >> 1) == 0 ? n[1] : n[0]) : 1, npop = d < 4 ? (dir >= 8 ? 1 : 3) : 1;
        M.dir = dir;
        // what I send towards direction `dir` is received from the neighbour on the opposite side
        const int ox = d < 4 ? (d == 0) - (d == 1) : h_t2_ex[d + 1], oy = d < 4 ? (d == 2) - (d == 3) : h_t2_ey[d + 1];
        M.send_to = t2_cart_rank(dims, c0 + ox, c1 + oy);
        M.recv_from = t2_cart_rank(dims, c0 - ox, c1 - oy);
        // face messages span the sender's interior range; both ends share that extent along the face (same coordinate there)
        M.send_count = M.send_to >= 0 ? (long long)n1 * npop : 0;
        M.recv_count = M.recv_from >= 0 ? (long long)n1 * npop : 0;
    }
}
extern "C" int mglc_t2d_msg_table(int total_nx, int total_ny, const int dims[2], int rank, mglc_halo_msg out[12]) {
    if (!dims || !out || dims[0] < 1 || dims[1] < 1 || rank < 0 || rank >= dims[0] * dims[1]) return MGLC_E_INVALID;
    const int c0 = rank / dims[1], c1 = rank % dims[1];
    int n[2], start;
    mglc_decompose_1d(total_nx, c0, dims[0], &n[0], &start);
    mglc_decompose_1d(total_ny, c1, dims[1], &n[1], &start);
    Msg msgs[T2_NMSG];
    memset(msgs, 0, sizeof msgs);
    t2_build_msgs(dims, n, c0, c1, msgs);
    for (int d = 0; d < T2_NMSG; ++d) {
        memset(&out[d], 0, sizeof out[d]);
        out[d].dir = msgs[d].dir; out[d].send_to = msgs[d].send_to; out[d].recv_from = msgs[d].recv_from;
        out[d].send_count = (int)msgs[d].send_count; out[d].recv_count = (int)msgs[d].recv_count;
        out[d].npop = (d >= 8 || (d >= 4 && d < 8)) ? 1 : 3;
    }
    return MGLC_OK;
}

static int t2_make_sub(mglc_t2d *h, int rank, int device, T2Sub **out) {
    T2Sub *S = new T2Sub();
    memset(S, 0, sizeof *S);
    S->device = device;
    S->coords[0] = rank / h->dims[1]; S->coords[1] = rank % h->dims[1];
    const int gn[2] = {h->d.total_nx, h->d.total_ny};
    for (int d = 0; d < 2; ++d) {
        if (gn[d] < h->dims[d]) { set_error("mglc_t2d_create: fewer cells than ranks along dim %d", d); delete S; return MGLC_E_INVALID; }
        mglc_decompose_1d(gn[d], S->coords[d], h->dims[d], &S->n[d], &S->start[d]);
    }
    const int c0 = S->coords[0], c1 = S->coords[1];
    S->nbr[0] = t2_cart_rank(h->dims, c0 + 1, c1); S->nbr[1] = t2_cart_rank(h->dims, c0 - 1, c1);
    S->nbr[2] = t2_cart_rank(h->dims, c0, c1 + 1); S->nbr[3] = t2_cart_rank(h->dims, c0, c1 - 1);
    for (int a = 5; a < 9; ++a) S->cnr[a - 5] = t2_cart_rank(h->dims, c0 + h_t2_ex[a], c1 + h_t2_ey[a]);
    S->g = make_geom2(S->n[0], S->n[1]);
    S->g.wall[0] = c0 == h->dims[0] - 1 && !h->p.perx; S->g.wall[1] = c0 == 0 && !h->p.perx;
    S->g.wall[2] = c1 == h->dims[1] - 1; S->g.wall[3] = c1 == 0;
    S->p = h->p;
    for (int d = 0; d < 2; ++d) { S->p.start[d] = S->start[d]; S->p.total[d] = gn[d]; }
    auto fail = [&](int rc) { t2_free_sub(S); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return fail(MGLC_E_CUDA); }
    if (cudaStreamCreateWithFlags(&S->s, cudaStreamNonBlocking) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaEventCreateWithFlags(&S->ev_packed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&S->ev_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&S->ev_t0) != cudaSuccess || cudaEventCreate(&S->ev_t1) != cudaSuccess) return fail(MGLC_E_CUDA);
    const size_t lf = 9 * (size_t)S->g.sq * sizeof(double), lg = 5 * (size_t)S->g.sq * sizeof(double);
    const size_t fld = (size_t)S->n[0] * S->n[1] * sizeof(double);
    struct { double **p; size_t bytes; } bufs[] = {
        {&S->F, lf}, {&S->P[0], lf}, {&S->P[1], lf}, {&S->G, lg}, {&S->Q[0], lg}, {&S->Q[1], lg},
        {&S->rho, fld}, {&S->u, fld}, {&S->v, fld}, {&S->T, fld}, {&S->up, fld}, {&S->vp, fld}, {&S->Tp, fld}, {&S->Fx, fld}, {&S->Fy, fld},
        {&S->stage, 9 * (size_t)(S->n[0] + 2) * (S->n[1] + 2) * sizeof(double)},
        {&S->scratch, (size_t)(T2_RED_WORDS + T2_RED_WORDS * T2_RED_BLOCKS) * sizeof(double)}};
    for (auto &b : bufs) {
        if (cudaMalloc((void **)b.p, b.bytes) != cudaSuccess) { (void)cudaGetLastError(); set_error("mglc_t2d_create: out of device memory"); return fail(MGLC_E_NOMEM); }
        cudaMemsetAsync(*b.p, 0, b.bytes, S->s);       // f_post = g_post = 0, initial.F90:334-335
    }
    t2_build_msgs(h->dims, S->n, c0, c1, S->msgs);
    for (int dir = 0; dir < T2_NMSG; ++dir) {
        Msg &M = S->msgs[dir];
        if (M.send_count && cudaMalloc((void **)&M.sbuf, M.send_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
        if (M.recv_count && cudaMalloc((void **)&M.rbuf, M.recv_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
    }
    {   // the message table must be the one the host-only plan publishes (mglc_halo_plan_2d, checked on the CPU against the oracle)
        mglc_halo_msg plan[12];
        int np = 0;
        if (mglc_halo_plan_2d(gn[0], gn[1], h->dims, rank, plan, &np) != MGLC_OK) return fail(MGLC_E_INVALID);
        for (int dir = 0; dir < T2_NMSG; ++dir)
            if (plan[dir].send_to != S->msgs[dir].send_to || plan[dir].recv_from != S->msgs[dir].recv_from ||
                plan[dir].send_count != S->msgs[dir].send_count || plan[dir].recv_count != S->msgs[dir].recv_count) {
                set_error("mglc_t2d_create: message %d differs from mglc_halo_plan_2d", dir);
                return fail(MGLC_E_STATE);
            }
    }
    if (cudaStreamSynchronize(S->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    *out = S;
    return MGLC_OK;
}

static int t2_new(mglc_t2d **out, const mglc_t2d_desc *d, const int dims_or_zero[2], int nranks) {
    if (!out || !d || nranks < 1) { set_error("mglc_t2d_create: bad arguments"); return MGLC_E_INVALID; }
    if (d->total_nx < 2 || d->total_ny < 2 || (d->arith != MGLC_ARITH_FAST && d->arith != MGLC_ARITH_STRICT) || !(d->Rayleigh > 0.0) ||
        !(d->Prandtl > 0.0) || !(d->Mach > 0.0)) {
        set_error("mglc_t2d_create: bad descriptor (%d x %d, arith %d, Ra %g, Pr %g, Ma %g)", d->total_nx, d->total_ny, d->arith, d->Rayleigh,
                  d->Prandtl, d->Mach);
        return MGLC_E_INVALID;
    }
    for (int f = 0; f < 4; ++f)
        if (d->bcT[f] < MGLC_BCT_ADIABATIC || d->bcT[f] > MGLC_BCT_PERIODIC) { set_error("mglc_t2d_create: bcT[%d]=%d", f, d->bcT[f]); return MGLC_E_INVALID; }
    const bool perx = d->bcT[0] == MGLC_BCT_PERIODIC || d->bcT[1] == MGLC_BCT_PERIODIC;
    if ((perx && d->bcT[0] != d->bcT[1]) || d->bcT[2] == MGLC_BCT_PERIODIC || d->bcT[3] == MGLC_BCT_PERIODIC) {
        set_error("mglc_t2d_create: MGLC_BCT_PERIODIC applies to both vertical walls together (seq/bouyancy2d_acc.F90:15,22)");
        return MGLC_E_INVALID;
    }
    if ((d->variant != MGLC_T2D_MPI && d->variant != MGLC_T2D_ACC) || d->lengthUnit < 0.0) {
        set_error("mglc_t2d_create: variant=%d lengthUnit=%g", d->variant, d->lengthUnit);
        return MGLC_E_INVALID;
    }
    MGLC_TRY(require_gpu());
    mglc_t2d *h = new mglc_t2d();
    h->d = *d; h->nranks = nranks; h->comm = nullptr;
    if (dims_or_zero && dims_or_zero[0] > 0) { h->dims[0] = dims_or_zero[0]; h->dims[1] = dims_or_zero[1]; }
    else { int d3[3]; mglc_dims_create_nd(nranks, 2, d3); h->dims[0] = d3[0]; h->dims[1] = d3[1]; }      // main.F90:22
    if (h->dims[0] * h->dims[1] != nranks) {
        set_error("mglc_t2d_create: dims %dx%d do not fit %d ranks", h->dims[0], h->dims[1], nranks);
        delete h;
        return MGLC_E_INVALID;
    }
    if (perx && h->dims[0] != 1) {
        set_error("mglc_t2d_create: periodic vertical walls need an undivided x direction (dims %dx%d); split along y only", h->dims[0], h->dims[1]);
        delete h;
        return MGLC_E_INVALID;
    }
    // module.F90:29,69-81 -- the same products in the same order
    h->lengthUnit = d->lengthUnit > 0.0 ? d->lengthUnit : (double)d->total_ny;
    h->tauf = 0.5 + d->Mach * h->lengthUnit * sqrt(3.0 * d->Prandtl / d->Rayleigh);
    h->viscosity = (h->tauf - 0.5) / 3.0;
    h->diffusivity = h->viscosity / d->Prandtl;
    T2Params &p = h->p;
    p.paraA = 20.0 * sqrt(3.0) * h->diffusivity - 4.0;
    const double gBeta1 = d->Rayleigh * h->viscosity * h->diffusivity / h->lengthUnit;
    p.gBeta = gBeta1 / h->lengthUnit / h->lengthUnit;
    p.Snu = 1.0 / h->tauf;
    p.Sq = 8.0 * (2.0 * h->tauf - 1.0) / (8.0 * h->tauf - 1.0);
    p.Qd = 3.0 - sqrt(3.0);
    p.Qnu = 4.0 * sqrt(3.0) - 6.0;
    p.Tref = d->Tref; p.rho0 = d->rho0; p.Thot = d->Thot; p.Tcold = d->Tcold;
    if (p.paraA >= 1.0 || p.paraA <= -4.0) {                                    // initial.F90:30-37: the reference stops here
        set_error("mglc_t2d_create: paraA = %g outside (-4, 1): reduce the Mach number (initial.F90:30)", p.paraA);
        delete h;
        return MGLC_E_INVALID;
    }
    p.perx = perx; p.variant = d->variant;
    for (int f = 0; f < 4; ++f) {
        p.bcT[f] = d->bcT[f] == MGLC_BCT_PERIODIC ? 0 : d->bcT[f];
        p.wallT[f] = (4.0 + p.paraA) / 10.0 * (d->bcT[f] == MGLC_BCT_CONST_HOT ? d->Thot : d->Tcold);    // evolution_g.F90:100,107,132,139
    }
    p.moving = 0;
    for (int q = 0; q < 8; ++q) { p.Uwall[q] = d->Uwall[q]; p.moving |= d->Uwall[q] != 0.0; }     // seq/R_B_2d.F90:118-120
    p.cornersT = d->cornersT != 0 && !perx;                                                        // :1086 #ifndef VerticalWallsPeriodicalT
    if (p.moving && perx) {
        set_error("mglc_t2d_create: moving walls with periodic vertical walls are not a configuration of the reference");
        delete h;
        return MGLC_E_INVALID;
    }
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_t2d_create(mglc_t2d **out, const mglc_t2d_desc *d, const int dims_or_zero[2], int nranks, int rank, int device,
                               mglc_comm *comm_or_null) {
    if (nranks > 1 && !comm_or_null) { set_error("mglc_t2d_create: %d ranks need a communicator (or use mglc_t2d_create_local)", nranks); return MGLC_E_INVALID; }
    if (rank < 0 || rank >= nranks) { set_error("mglc_t2d_create: rank=%d of %d", rank, nranks); return MGLC_E_INVALID; }
    mglc_t2d *h = nullptr;
    MGLC_TRY(t2_new(&h, d, dims_or_zero, nranks));
    h->comm = comm_or_null;
    T2Sub *S = nullptr;
    int rc = t2_make_sub(h, rank, device, &S);
    if (rc) { delete h; return rc; }
    h->subs.push_back(S);
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_t2d_create_local(mglc_t2d **out, const mglc_t2d_desc *d, const int dims_or_zero[2], int nranks,
                                     const int *devices_or_null) {
    mglc_t2d *h = nullptr;
    MGLC_TRY(t2_new(&h, d, dims_or_zero, nranks));
    for (int r = 0; r < nranks; ++r) {
        T2Sub *S = nullptr;
        int rc = t2_make_sub(h, r, devices_or_null ? devices_or_null[r] : 0, &S);
        if (rc) { mglc_t2d_destroy(h); return rc; }
        h->subs.push_back(S);
    }
    for (T2Sub *a : h->subs)
        for (T2Sub *b : h->subs)
            if (a->device != b->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->device, b->device);
                if (can) { cudaSetDevice(a->device); cudaDeviceEnablePeerAccess(b->device, 0); (void)cudaGetLastError(); }
            }
    for (T2Sub *S : h->subs) h->ports.push_back(Port{S->device, S->s, S->ev_packed, S->ev_copied, S->msgs, T2_NMSG});
    *out = h;
    return MGLC_OK;
}

static int t2_sub(mglc_t2d *h, int r, T2Sub **S) {
    if (!h || r < 0 || r >= (int)h->subs.size()) { set_error("mglc_t2d: bad handle or local index %d", r); return MGLC_E_INVALID; }
    *S = h->subs[r];
    return t2_use(*S);
}

extern "C" int mglc_t2d_nlocal(mglc_t2d *h, int *n) {
    if (!h || !n) return MGLC_E_INVALID;
    *n = (int)h->subs.size();
    return MGLC_OK;
}
extern "C" int mglc_t2d_info(mglc_t2d *h, int r, int dims[2], int ln[2], int start[2], int coords[2], int nbr[8]) {
    if (!h || r < 0 || r >= (int)h->subs.size()) return MGLC_E_INVALID;
    T2Sub *S = h->subs[r];
    if (dims) memcpy(dims, h->dims, 8);
    if (ln) memcpy(ln, S->n, 8);
    if (start) memcpy(start, S->start, 8);
    if (coords) memcpy(coords, S->coords, 8);
    if (nbr) { memcpy(nbr, S->nbr, 16); memcpy(nbr + 4, S->cnr, 16); }
    return MGLC_OK;
}
extern "C" int mglc_t2d_params(mglc_t2d *h, double out[10]) {
    if (!h || !out) return MGLC_E_INVALID;
    const double v[10] = {h->tauf, h->viscosity, h->diffusivity, h->p.paraA, h->p.gBeta, h->p.Snu, h->p.Sq, h->p.Qd, h->p.Qnu, h->lengthUnit};
    memcpy(out, v, sizeof v);
    return MGLC_OK;
}

static dim3 t2_grid_h(const T2Sub *S, int halo = 0) { return dim3((S->n[0] + 2 * halo + 127) / 128, S->n[1] + 2 * halo); }

static int t2_put_lattice(T2Sub *S, int nq, const double *host, double *dev, int with_halo) {
    if (!host) return MGLC_OK;
    const size_t cells = with_halo ? (size_t)(S->n[0] + 2) * (S->n[1] + 2) : (size_t)S->n[0] * S->n[1];
    MGLC_CUDA(cudaMemcpyAsync(S->stage, host, nq * cells * sizeof(double), cudaMemcpyHostToDevice, S->s));
    k_t2_aos_to_soa<<<t2_grid_h(S, with_halo), 128, 0, S->s>>>(S->g, nq, S->stage, dev, with_halo);
    S->launches += 1;
    MGLC_CUDA(cudaStreamSynchronize(S->s));       // the staging buffer is reused and `host` may be pageable
    return MGLC_OK;
}
static int t2_get_lattice(T2Sub *S, int nq, double *host, const double *dev, int with_halo) {
    if (!host) return MGLC_OK;
    const size_t cells = with_halo ? (size_t)(S->n[0] + 2) * (S->n[1] + 2) : (size_t)S->n[0] * S->n[1];
    k_t2_soa_to_aos<<<t2_grid_h(S, with_halo), 128, 0, S->s>>>(S->g, nq, dev, S->stage, with_halo);
    S->launches += 1;
    MGLC_CUDA(cudaMemcpyAsync(host, S->stage, nq * cells * sizeof(double), cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}

// fields[] order: rho, u, v, T, Fx, Fy
extern "C" int mglc_t2d_upload(mglc_t2d *h, int r, const double *f, const double *f_post, const double *g, const double *g_post,
                               const double *const fields_or_null[6]) {
    T2Sub *S;
    MGLC_TRY(t2_sub(h, r, &S));
    MGLC_TRY(t2_put_lattice(S, 9, f, S->F, 0));
    MGLC_TRY(t2_put_lattice(S, 9, f_post, S->P[S->cur], 1));
    MGLC_TRY(t2_put_lattice(S, 5, g, S->G, 0));
    MGLC_TRY(t2_put_lattice(S, 5, g_post, S->Q[S->cur], 1));
    const size_t fld = (size_t)S->n[0] * S->n[1] * sizeof(double);
    double *dev[6] = {S->rho, S->u, S->v, S->T, S->Fx, S->Fy};
    for (int q = 0; fields_or_null && q < 6; ++q)
        if (fields_or_null[q]) MGLC_CUDA(cudaMemcpyAsync(dev[q], fields_or_null[q], fld, cudaMemcpyHostToDevice, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
extern "C" int mglc_t2d_download(mglc_t2d *h, int r, double *f, double *f_post, double *g, double *g_post, double *const fields_or_null[6]) {
    T2Sub *S;
    MGLC_TRY(t2_sub(h, r, &S));
    MGLC_TRY(t2_get_lattice(S, 9, f, S->F, 0));
    MGLC_TRY(t2_get_lattice(S, 9, f_post, S->P[S->cur], 1));
    MGLC_TRY(t2_get_lattice(S, 5, g, S->G, 0));
    MGLC_TRY(t2_get_lattice(S, 5, g_post, S->Q[S->cur], 1));
    const size_t fld = (size_t)S->n[0] * S->n[1] * sizeof(double);
    double *dev[6] = {S->rho, S->u, S->v, S->T, S->Fx, S->Fy};
    for (int q = 0; fields_or_null && q < 6; ++q)
        if (fields_or_null[q]) MGLC_CUDA(cudaMemcpyAsync(fields_or_null[q], dev[q], fld, cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    MGLC_CUDA(cudaGetLastError());
    return MGLC_OK;
}

extern "C" int mglc_t2d_initial(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    // #ifdef VerticalWallsConstT: T linear in x (initial.F90:252-261); #ifdef HorizontalWallsConstT: linear in y, applied after (:262-271)
    auto constT = [&](int f) { return h->d.bcT[f] == MGLC_BCT_CONST_HOT || h->d.bcT[f] == MGLC_BCT_CONST_COLD; };
    const bool vertT = constT(0) || constT(1), horT = constT(2) || constT(3);
    const int profile = horT ? 2 : vertT ? 1 : 0;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        const int axis = profile == 2 ? 1 : 0;
        k_t2_initial<<<t2_grid_h(S), 128, 0, S->s>>>(S->g, S->p, profile, S->start[axis], axis ? h->d.total_ny : h->d.total_nx, S->F, S->G,
                                                     S->rho, S->u, S->v, S->T, S->up, S->vp, S->Tp);
        MGLC_CUDA(cudaMemsetAsync(S->P[S->cur], 0, 9 * (size_t)S->g.sq * sizeof(double), S->s));     // f_post = 0, initial.F90:334
        MGLC_CUDA(cudaMemsetAsync(S->Q[S->cur], 0, 5 * (size_t)S->g.sq * sizeof(double), S->s));     // g_post = 0, :335
        S->launches += 1;
    }
    return MGLC_OK;
}

static int t2_collision(mglc_t2d *h) {
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        S->launches += (h->d.arith == MGLC_ARITH_STRICT ? strict::launch_t2_collision : fast::launch_t2_collision)(
            S->g, S->p, S->F, S->rho, S->u, S->v, S->T, S->P[S->cur], S->Fx, S->Fy, S->s);
    }
    return MGLC_OK;
}
static int t2_collisionT(mglc_t2d *h) {
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        S->launches += (h->d.arith == MGLC_ARITH_STRICT ? strict::launch_t2_collisionT : fast::launch_t2_collisionT)(
            S->g, S->p, S->G, S->u, S->v, S->T, S->Q[S->cur], S->s);
    }
    return MGLC_OK;
}

static int t2_pack(T2Sub *S, cudaStream_t s) {
    for (int dir = 0; dir < T2_NMSG; ++dir) {
        Msg &M = S->msgs[dir];
        if (!M.send_count || M.skip) continue;
        int n1, npop;
        t2_msg_dims(S, dir, n1, npop);
        k_t2_pack<<<(unsigned)((M.send_count + 127) / 128), 128, 0, s>>>(S->g, dir >= 8 ? S->Q[S->cur] : S->P[S->cur], dir, n1, npop, M.sbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}
static int t2_unpack(T2Sub *S, cudaStream_t s) {
    for (int dir = 0; dir < T2_NMSG; ++dir) {
        Msg &M = S->msgs[dir];
        if (!M.recv_count || M.skip) continue;
        int n1, npop;
        t2_msg_dims(S, dir, n1, npop);
        k_t2_unpack<<<(unsigned)((M.recv_count + 127) / 128), 128, 0, s>>>(S->g, dir >= 8 ? S->Q[S->cur] : S->P[S->cur], dir, n1, npop, M.rbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}

// message_passing_f() / message_passing_g(), message_exchange.F90: the 12 + 4 (+ 4) MPI_Sendrecv become one grouped NCCL
// operation (or device-to-device copies between the subdomains of one process).  which: 1 = f, 2 = g, 3 = both (the fused step)
static int t2_exchange(mglc_t2d *h, int which) {
    if (h->nranks == 1) return MGLC_OK;
    for (T2Sub *S : h->subs)
        for (int dir = 0; dir < T2_NMSG; ++dir) S->msgs[dir].skip = !(which & (dir >= 8 ? 2 : 1));
    if (h->comm) {
        T2Sub *S = h->subs[0];
        MGLC_TRY(t2_use(S));
        MGLC_TRY(t2_pack(S, S->s));
        MGLC_TRY(halo_nccl_sendrecv(S->msgs, T2_NMSG, h->comm, S->s));
        MGLC_TRY(t2_unpack(S, S->s));
        return MGLC_OK;
    }
    return halo_local_exchange(
        h->ports, [&](int r, cudaStream_t s) { return t2_pack(h->subs[r], s); }, [&](int r, cudaStream_t s) { return t2_unpack(h->subs[r], s); });
}

extern "C" int mglc_t2d_collision(mglc_t2d *h) { if (!h) return MGLC_E_INVALID; return t2_collision(h); }
extern "C" int mglc_t2d_collisionT(mglc_t2d *h) { if (!h) return MGLC_E_INVALID; return t2_collisionT(h); }
extern "C" int mglc_t2d_exchange_f(mglc_t2d *h) { if (!h) return MGLC_E_INVALID; return t2_exchange(h, 1); }
extern "C" int mglc_t2d_exchange_g(mglc_t2d *h) { if (!h) return MGLC_E_INVALID; return t2_exchange(h, 2); }
extern "C" int mglc_t2d_streaming(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_streaming<<<t2_grid_h(S), 128, 0, S->s>>>(S->g, 9, S->P[S->cur], S->F);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_streamingT(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_streaming<<<t2_grid_h(S), 128, 0, S->s>>>(S->g, 5, S->Q[S->cur], S->G);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_bounceback(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        const int cells = 2 * S->n[0] + 2 * std::max(S->n[1] - 2, 0);
        k_t2_bounceback<<<(cells + 127) / 128, 128, 0, S->s>>>(S->g, S->p, S->P[S->cur], S->F, S->rho);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_bouncebackT(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        const int cells = 2 * S->n[0] + 2 * std::max(S->n[1] - 2, 0);
        k_t2_bouncebackT<<<(cells + 127) / 128, 128, 0, S->s>>>(S->g, S->p, S->Q[S->cur], S->G);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_macro(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_macro<<<t2_grid_h(S), 128, 0, S->s>>>(S->g, S->F, S->Fx, S->Fy, S->rho, S->u, S->v);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_macroT(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_macroT<<<t2_grid_h(S), 128, 0, S->s>>>(S->g, S->G, S->T);
        S->launches += 1;
    }
    return MGLC_OK;
}

// nsteps loop bodies (main.F90:84-108), rotated by half a step: collision() + collisionT() once, then nsteps-1 times
// [exchange f_post and g_post -> the ten remaining subroutines of this body and the two collisions of the next in one kernel],
// then exchange -> streaming + bounceback + streamingT + bouncebackT + macro + macroT.  f, g, f_post, g_post (interior +
// exchanged halos), rho, u, v, T, Fx, Fy afterwards are the reference's after the same number of iterations.
static int t2_step_impl(mglc_t2d *h, int nsteps) {
    if (nsteps < 0) { set_error("mglc_t2d_step: nsteps=%d", nsteps); return MGLC_E_INVALID; }
    if (nsteps == 0) return MGLC_OK;
    const bool strict_build = h->d.arith == MGLC_ARITH_STRICT;
    MGLC_TRY(t2_collision(h));
    MGLC_TRY(t2_collisionT(h));
    int it = 1;
    if (h->nranks == 1 && (long long)h->subs[0]->n[0] * h->subs[0]->n[1] <= t2_graph_max_cells()) {
        T2Sub *S = h->subs[0];
        MGLC_TRY(t2_use(S));
        auto fused = strict_build ? strict::launch_t2_fused : fast::launch_t2_fused;
        while (nsteps - it >= T2_GRAPH_STEPS) {
            cudaGraphExec_t &ge = S->gexec[S->cur];
            if (!ge) {
                cudaGraph_t gr = nullptr;
                MGLC_CUDA(cudaStreamBeginCapture(S->s, cudaStreamCaptureModeRelaxed));
                int c = S->cur;
                for (int q = 0; q < T2_GRAPH_STEPS; ++q, c ^= 1) fused(S->g, S->p, S->P[c], S->P[c ^ 1], S->Q[c], S->Q[c ^ 1], S->Fy, S->rho, S->s);
                const cudaError_t ce = cudaStreamEndCapture(S->s, &gr);
                if (ce != cudaSuccess) { if (gr) cudaGraphDestroy(gr); (void)cudaGetLastError(); set_error("mglc_t2d_step: graph capture failed: %s", cudaGetErrorString(ce)); return MGLC_E_CUDA; }
                const cudaError_t ie = cudaGraphInstantiate(&ge, gr, 0);
                cudaGraphDestroy(gr);
                if (ie != cudaSuccess) { ge = nullptr; set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ie)); return MGLC_E_CUDA; }
            }
            MGLC_CUDA(cudaGraphLaunch(ge, S->s));
            S->launches += T2_GRAPH_STEPS;
            it += T2_GRAPH_STEPS;
        }
    }
    for (; it < nsteps; ++it) {
        MGLC_TRY(t2_exchange(h, 3));
        for (T2Sub *S : h->subs) {
            MGLC_TRY(t2_use(S));
            S->launches += (strict_build ? strict::launch_t2_fused : fast::launch_t2_fused)(S->g, S->p, S->P[S->cur], S->P[S->cur ^ 1], S->Q[S->cur],
                                                                                            S->Q[S->cur ^ 1], S->Fy, S->rho, S->s);
            S->cur ^= 1;
        }
    }
    MGLC_TRY(t2_exchange(h, 3));
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        S->launches += strict::launch_t2_stream_macro(S->g, S->p, S->P[S->cur], S->F, S->Q[S->cur], S->G, S->Fy, S->rho, S->u, S->v, S->T, S->s);
    }
    return MGLC_OK;
}
extern "C" int mglc_t2d_step(mglc_t2d *h, int nsteps) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(t2_step_impl(h, nsteps));
    for (T2Sub *S : h->subs) { MGLC_TRY(t2_use(S)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
extern "C" int mglc_t2d_step_timed(mglc_t2d *h, int nsteps, float *ms) {
    if (!h || !ms) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) { MGLC_TRY(t2_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    for (T2Sub *S : h->subs) { MGLC_TRY(t2_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t0, S->s)); }
    MGLC_TRY(t2_step_impl(h, nsteps));
    for (T2Sub *S : h->subs) { MGLC_TRY(t2_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t1, S->s)); }
    float worst = 0.f;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        MGLC_CUDA(cudaEventSynchronize(S->ev_t1));
        MGLC_CUDA(cudaGetLastError());
        float t = 0.f;
        MGLC_CUDA(cudaEventElapsedTime(&t, S->ev_t0, S->ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MGLC_OK;
}

// reduce nwords partial sums per subdomain, then over the subdomains in rank order (and over the ranks with NCCL)
static int t2_reduce(mglc_t2d *h, int nwords, double *tot) {
    for (int q = 0; q < nwords; ++q) tot[q] = 0.0;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_reduce_final<<<1, 1, 0, S->s>>>(T2_RED_BLOCKS, nwords, S->scratch);
        S->launches += 1;
        if (h->comm && h->nranks > 1) MGLC_NCCL(ncclAllReduce(S->scratch, S->scratch, nwords, ncclDouble, ncclSum, h->comm->nccl, S->s));
        double e[T2_RED_WORDS];
        MGLC_CUDA(cudaMemcpyAsync(e, S->scratch, nwords * sizeof(double), cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        MGLC_CUDA(cudaGetLastError());
        for (int q = 0; q < nwords; ++q) tot[q] += e[q];
    }
    return MGLC_OK;
}

// check(): check.F90:1-51 (rank sums + 4 MPI_Allreduce)
extern "C" int mglc_t2d_check(mglc_t2d *h, double *errorU, double *errorT) {
    if (!h || !errorU || !errorT) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_check_partial<<<T2_RED_BLOCKS, 256, 0, S->s>>>((long long)S->n[0] * S->n[1], S->u, S->v, S->T, S->up, S->vp, S->Tp, S->scratch);
        S->launches += 1;
    }
    double t[4];
    MGLC_TRY(t2_reduce(h, 4, t));
    *errorU = sqrt(t[0]) / sqrt(t[1]);
    *errorT = t[2] / t[3];
    return MGLC_OK;
}

// the volume averages of calNuRe(), NuRe.F90:27-78: out = angular momentum / N, NuVolAvg, ReVolAvg   (N = total_nx*total_ny)
extern "C" int mglc_t2d_nure(mglc_t2d *h, double out[3]) {
    if (!h || !out) return MGLC_E_INVALID;
    const int nxHalf = (h->d.total_nx - 1) / 2 + 1, nyHalf = (h->d.total_ny - 1) / 2 + 1;       // module.F90:57
    for (T2Sub *S : h->subs) {
        MGLC_TRY(t2_use(S));
        k_t2_nure_partial<<<T2_RED_BLOCKS, 256, 0, S->s>>>(S->n[0], S->n[1], S->start[0], S->start[1], nxHalf, nyHalf, S->u, S->v, S->T, S->scratch);
        S->launches += 1;
    }
    double t[3];
    MGLC_TRY(t2_reduce(h, 3, t));
    const double fluidNumMax = (double)((long long)h->d.total_nx * h->d.total_ny);             // initial.F90:16
    out[0] = t[0] / fluidNumMax;
    out[1] = t[1] / fluidNumMax * h->lengthUnit / h->diffusivity + 1.0;
    out[2] = sqrt(t[2] / fluidNumMax) * h->lengthUnit / h->viscosity;
    return MGLC_OK;
}

extern "C" int mglc_t2d_launch_count(mglc_t2d *h, long long *n) {
    if (!h || !n) return MGLC_E_INVALID;
    long long t = 0;
    for (T2Sub *S : h->subs) t += S->launches;
    *n = t;
    return MGLC_OK;
}
extern "C" int mglc_t2d_sync(mglc_t2d *h) {
    if (!h) return MGLC_E_INVALID;
    for (T2Sub *S : h->subs) { MGLC_TRY(t2_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
