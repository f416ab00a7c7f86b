// halo.cuh -- internal: the halo-message plumbing shared by every solver behind the C ABI.
// A message is one packed device buffer per neighbour and direction (the replacement for the
// reference's MPI derived datatypes + MPI_Sendrecv, e.g. L3/ex_sendrecv.f90, LAP:223-254).  Two
// transports move it: NCCL send/recv grouped per step (one process per GPU) or event-ordered
// device-to-device copies between subdomains owned by one process (tests, single-process multi-GPU).
#pragma once
#include <nccl.h>

#include <vector>

#include "common.cuh"

struct mglc_comm {
    ncclComm_t nccl;
    int nranks, rank, device;
};

namespace mglc {

struct Msg {
    int dir, send_to, recv_from;
    long long send_count, recv_count;
    double *sbuf, *rbuf;
    int skip;            // 1 = leave this message out of the current exchange (e.g. g messages during f's)
};

#define MGLC_NCCL(call)                                                                        \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            ::mglc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
            return MGLC_E_NCCL;                                                                \
        }                                                                                      \
    } while (0)
#define MGLC_TRY(call)                  \
    do {                                \
        int rc_ = (call);               \
        if (rc_ != MGLC_OK) return rc_; \
    } while (0)

// all of a subdomain's messages as ONE grouped NCCL operation on stream s
inline int halo_nccl_sendrecv(const Msg *msgs, int n, mglc_comm *comm, cudaStream_t s) {
    MGLC_NCCL(ncclGroupStart());
    for (int m = 0; m < n; ++m) {
        const Msg &M = msgs[m];
        if (M.skip) continue;
        if (M.send_count) MGLC_NCCL(ncclSend(M.sbuf, (size_t)M.send_count, ncclDouble, M.send_to, comm->nccl, s));
        if (M.recv_count) MGLC_NCCL(ncclRecv(M.rbuf, (size_t)M.recv_count, ncclDouble, M.recv_from, comm->nccl, s));
    }
    MGLC_NCCL(ncclGroupEnd());
    return MGLC_OK;
}

// one subdomain's end of the in-process transport; message m of the sender pairs with message m of
// the receiver (same direction index on both sides)
struct Port {
    int device;
    cudaStream_t s;
    cudaEvent_t ev_packed, ev_copied;
    Msg *msgs;
    int nmsgs;
};

// pack(r, stream) / unpack(r, stream) launch rank r's pack / unpack kernels and return MGLC_OK or an error
template <class Pack, class Unpack>
int halo_local_exchange(std::vector<Port> &ports, Pack pack, Unpack unpack) {
    const int P = (int)ports.size();
    for (int r = 0; r < P; ++r) {
        Port &p = ports[r];
        MGLC_CUDA(cudaSetDevice(p.device));
        // the peers that copied out of my send buffers last time must be done before I overwrite them
        for (int m = 0; m < p.nmsgs; ++m)
            if (p.msgs[m].send_count && !p.msgs[m].skip)
                MGLC_CUDA(cudaStreamWaitEvent(p.s, ports[p.msgs[m].send_to].ev_copied, 0));
        MGLC_TRY(pack(r, p.s));
        MGLC_CUDA(cudaEventRecord(p.ev_packed, p.s));
    }
    for (int r = 0; r < P; ++r) {
        Port &p = ports[r];
        MGLC_CUDA(cudaSetDevice(p.device));
        for (int m = 0; m < p.nmsgs; ++m) {
            Msg &M = p.msgs[m];
            if (!M.recv_count || M.skip) continue;
            Port &src = ports[M.recv_from];
            MGLC_CUDA(cudaStreamWaitEvent(p.s, src.ev_packed, 0));
            MGLC_CUDA(cudaMemcpyPeerAsync(M.rbuf, p.device, src.msgs[m].sbuf, src.device,
                                          (size_t)M.recv_count * sizeof(double), p.s));
        }
        MGLC_CUDA(cudaEventRecord(p.ev_copied, p.s));
        MGLC_TRY(unpack(r, p.s));
    }
    return MGLC_OK;
}

inline int require_gpu() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available: libmglc.so has no CPU fallback (%s)",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MGLC_E_NOGPU;
    }
    return MGLC_OK;
}

}  // namespace mglc
