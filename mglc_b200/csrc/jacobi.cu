// jacobi.cu -- Jacobi halo-exchange path of the reference (MPI/Laplace/fortran/jacobi2d_mpi.f90, "LAP")
// and its 3-D extension (BASELINE.json config 2), behind the mglc_jacobi_* entry points of mglc.h.
//
// Device layout: A[k][j][x], one ghost layer, rows padded like the lattice (common.cuh Geom: interior
// cell i=1 sits 128-byte aligned).  2-D problems live in plane k = 1 of a 3-plane array.
// Kernels (all bit-defined: adds in the reference's order and one multiply by a constant, -fmad=false):
//   k_jacobi2d   A_new = 0.25 * (A(i-1,j) + A(i+1,j) + A(i,j-1) + A(i,j+1) + f)        LAP:170-182
//   k_jacobi3d   A_new = (1/6) * (x- + x+ + y- + y+ + z- + z+ + f); register blocking: a thread marches along z with
//                the z-1 / z / z+1 values of 4 consecutive rows in registers, x neighbours by warp shuffle, so a
//                cell costs one DRAM read and one write (16 B, 24 B with a source term)
//   k_jacobi3d_tma   the same update as a persistent TMA pipeline (the default in 3-D): one CTA per SM marches a tile of
//                128 x th cells (th <= 16) along z; a producer thread streams the halo'd planes (132 x (th+2) doubles) into a
//                ring of shared-memory stages with cp.async.bulk.tensor (completion on mbarriers), 16 consumer warps keep
//                z-1 / z / z+1 of their 4 cells in registers and take the x / y neighbours from the stage; the loads in
//                flight no longer depend on registers or occupancy
//   k_face_pack / k_face_unpack   replace the contiguous-row and MPI_Type_vector column messages of
//                exchange_message (LAP:223-254)
//   k_absdiff_max   check_diff (LAP:185-204)
#include <cuda.h>      // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time, libcuda is not linked)

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "halo.cuh"

using namespace mglc;

namespace {

struct JacSub {
    int n[3], coords[3], start[3], nbr[6];
    int device;
    Geom g;
    double *A[2];        // A[cur] = A, A[cur^1] = A_new
    int cur;
    double *f;           // source term, allocated on first upload (NULL = identically zero)
    double *A_p;         // previous-check copy
    double *scratch;     // check_diff result (one double, compared as an integer)
    cudaStream_t s;
    cudaEvent_t ev_packed, ev_copied, ev_t0, ev_t1;
    Msg msgs[6];
    long long launches;
    // direct halo stores (the fused step on several GPUs): the sweep writes its boundary values straight into the neighbours'
    // ghost layers of the array it is producing -- what exchange_message() (LAP:223-254) would bring them before the next
    // sweep -- and a flag barrier (k_halo_signal / k_halo_wait, shared with the lattice-Boltzmann path) orders the steps
    int direct, direct_valid, parity_sent;
    double *peerA[6][2];             // face f: the neighbour's A[0], A[1] as seen from this GPU
    Geom peer_g[6];
    unsigned long long *flags, epoch;
    int *d_err;
    SyncTable sync;
    cudaEvent_t ev_done[2];
    std::vector<void *> *ipc_opened;
    CUtensorMap tmap[2]; // 3-D tiled maps of A[0], A[1] (box TSX x TSY x 1 doubles) for k_jacobi3d_tma
    CUtensorMap *tmap_dev; // the same two maps in device memory (MGLC_JACOBI_TMAP=global)
    int tma_ok, tma_shape, tma_th, tma_chunks;   // CTA shape, tile height and z chunks of the TMA sweep (jac_tma_shape)
};

__device__ __forceinline__ void face_cell(const Geom &g, int face, int ghost, int t1, int t2, int &i, int &j, int &k) {
    const int axis = face >> 1, plus = !(face & 1);
    const int nfix = (axis == 0) ? g.nx : (axis == 1 ? g.ny : g.nz);
    const int fix = ghost ? (plus ? 0 : nfix + 1) : (plus ? nfix : 1);
    i = (axis == 0) ? fix : 1 + t1;
    j = (axis == 1) ? fix : (axis == 0 ? 1 + t1 : 1 + t2);
    k = (axis == 2) ? fix : 1 + t2;
}
// ndim == 2: the only plane is k = 1 and t2 == 0 for every face
__global__ void k_face_pack(Geom g, const double *__restrict__ A, int face, int n1, int n2, double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n1 * n2) return;
    int i, j, k;
    face_cell(g, face, 0, (int)(t % n1), (int)(t / n1), i, j, k);
    buf[t] = A[g.idx(0, i, j, k)];
}
__global__ void k_face_unpack(Geom g, double *__restrict__ A, int face, int n1, int n2, const double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n1 * n2) return;
    int i, j, k;
    face_cell(g, face, 1, (int)(t % n1), (int)(t / n1), i, j, k);
    A[g.idx(0, i, j, k)] = buf[t];
}

// where the boundary values of a sweep also go when the neighbours' arrays are mapped (direct halo stores)
struct JacPeers {
    unsigned mask;                   // bit f: a neighbour across face f (+x,-x,+y,-y,+z,-z) whose array is mapped
    double *B[6];                    // its A_new
    long long sy[6], sz[6];
    int n[6];                        // its extent along the face normal (the minus-side ghost sits at n + 1)
    const int *err;                  // sticky error word of the barrier: once set the sweep does nothing
};
// The store is issued as PTX without a memory clobber: nothing in these kernels ever reads a neighbour's array, so the compiler
// may keep hoisting the next planes' loads across it (as a plain C++ store through a non-restrict pointer it fenced them and the
// register-blocked sweep lost a third of its speed).
__device__ __forceinline__ void jac_remote_store(double *p, double val) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(val)); }
__device__ __forceinline__ void jac_peer_store(const JacPeers &P, const Geom &g, int i, int j, int k, double val) {
    const unsigned pm = P.mask;
    if (i == g.nx && (pm & 1u)) jac_remote_store(P.B[0] + (k * P.sz[0] + j * P.sy[0] + (OX - 1)), val);
    if (i == 1 && (pm & 2u)) jac_remote_store(P.B[1] + (k * P.sz[1] + j * P.sy[1] + (P.n[1] + OX)), val);
    if (j == g.ny && (pm & 4u)) jac_remote_store(P.B[2] + (k * P.sz[2] + (i + OX - 1)), val);
    if (j == 1 && (pm & 8u)) jac_remote_store(P.B[3] + (k * P.sz[3] + (P.n[3] + 1) * P.sy[3] + (i + OX - 1)), val);
    if (k == g.nz && (pm & 16u)) jac_remote_store(P.B[4] + (j * P.sy[4] + (i + OX - 1)), val);
    if (k == 1 && (pm & 32u)) jac_remote_store(P.B[5] + ((P.n[5] + 1) * P.sz[5] + j * P.sy[5] + (i + OX - 1)), val);
}

// The hand-over as its own small launch after the sweep: the six boundary faces of A_new go into the neighbours' ghost layers
// (blockIdx.y = face).  In-kernel stores (PEER = true below) turned the sweep's inner loop into mostly address arithmetic --
// with 4 CTAs across x half of all CTAs touch an x face -- and cost 0.24 ms of a 0.37 ms sweep; this launch moves the same
// 12.6 MB at 512^3 in a few microseconds (profiles/r2f_*).
__global__ void __launch_bounds__(256) k_jac_push(Geom g, int ndim, const double *__restrict__ B, JacPeers P) {
    const int face = blockIdx.y;
    if (!(P.mask >> face & 1u) || *P.err) return;
    const int axis = face >> 1;
    const int n1 = (axis == 0) ? g.ny : g.nx, n2 = (ndim == 2) ? 1 : ((axis == 2) ? g.ny : g.nz);
    const int nfix = (axis == 0) ? g.nx : (axis == 1 ? g.ny : g.nz), fix = (face & 1) ? 1 : nfix;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)n1 * n2; t += (long long)gridDim.x * blockDim.x) {
        const int t1 = 1 + (int)(t % n1), t2 = 1 + (int)(t / n1);
        const int i = (axis == 0) ? fix : t1;
        const int j = (axis == 1) ? fix : (axis == 0 ? t1 : t2);
        const int k = (ndim == 2) ? 1 : ((axis == 2) ? fix : t2);
        jac_peer_store(P, g, i, j, k, B[g.idx(0, i, j, k)]);
    }
}

template <bool HAS_F, bool PEER>
__global__ void __launch_bounds__(128) k_jacobi2d(Geom g, const double *__restrict__ A, const double *__restrict__ f,
                                                  double *__restrict__ B, JacPeers P) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > g.nx) return;
    const bool on_face = PEER && ((blockIdx.x == 0) | (blockIdx.x == gridDim.x - 1) | (j == 1) | (j == g.ny)) && *P.err == 0;
    const long long c = g.idx(0, i, j, 1);
    double s = A[c - 1] + A[c + 1];
    s += A[c - g.sy];
    s += A[c + g.sy];
    s += HAS_F ? f[c] : 0.0;
    const double val = 0.25 * s;
    B[c] = val;
    if (PEER && on_face) jac_peer_store(P, g, i, j, 1, val);
}

// Register blocking, no barriers: a thread owns JRY consecutive rows of one x column and marches `kch` planes along
// z with the z-1 / z / z+1 values of its JRY cells in registers.  y neighbours are the thread's own registers
// (plus one row above and one below the strip per plane), x neighbours come from the adjacent lanes by shuffle
// (warp-edge lanes read them), z neighbours from the marching registers: per cell one DRAM read, one write and
// about 2/JRY extra L2 reads.
constexpr int JTX = 128;
// PF: the loads of plane k+2 (and of the rows above / below the strip in plane k+1) are issued before plane k is computed, so a
// whole iteration of arithmetic and stores lies between a load and its first use
template <bool HAS_F, int JRY, bool PEER, bool PF>
__global__ void __launch_bounds__(JTX, JRY <= 4 ? (PF ? 7 : 8) : 4) k_jacobi3d(Geom g, const double *__restrict__ A, const double *__restrict__ f,
                                                                   double *__restrict__ B, int k_lo, int k_hi, int kch, JacPeers P) {
    const int i = 1 + blockIdx.x * JTX + threadIdx.x;
    const int j0 = 1 + blockIdx.y * JRY;
    const int k0 = k_lo + blockIdx.z * kch;
    const int k1 = min(k0 + kch - 1, k_hi);
    // only CTAs that touch a face of the block have boundary values to hand over (block-uniform test: the interior skips the
    // address arithmetic of the six possible messages); after a failed barrier nothing is stored into a neighbour
    const bool on_face = PEER && ((blockIdx.x == 0) | (blockIdx.x == gridDim.x - 1) | (j0 == 1) | (j0 + JRY - 1 >= g.ny) | (k0 == 1) | (k1 == g.nz)) &&
                         *P.err == 0;
    const int lane = threadIdx.x & 31;
    const bool active = i <= g.nx;
    const bool edge_m = lane == 0, edge_p = lane == 31 || i >= g.nx;
    const long long sy = g.sy, sz = g.sz;
    long long row[JRY];                           // rows past ny fold onto the ghost row ny+1: loaded, never stored
#pragma unroll
    for (int r = 0; r < JRY; ++r) row[r] = (long long)(min(j0 + r, g.ny + 1) - j0) * sy;
    const long long row_p = (long long)(min(j0 + JRY, g.ny + 1) - j0) * sy;
    long long c = g.idx(0, min(i, g.nx), j0, k0);
    double below[JRY], cen[JRY], up[JRY], nxt[JRY];
    double ym_n = 0.0, yp_n = 0.0;
#pragma unroll
    for (int r = 0; r < JRY; ++r) { below[r] = A[c + row[r] - sz]; cen[r] = A[c + row[r]]; }
    if (PF) {
#pragma unroll
        for (int r = 0; r < JRY; ++r) nxt[r] = A[c + row[r] + sz];
        ym_n = A[c - sy]; yp_n = A[c + row_p];
    }
    for (int k = k0; k <= k1; ++k) {
        double ym, yp;
        if (PF) {
#pragma unroll
            for (int r = 0; r < JRY; ++r) up[r] = nxt[r];
            ym = ym_n; yp = yp_n;
            if (k < k1) {                                   // plane k+2 and the strip's outer rows of plane k+1, for the next turn
#pragma unroll
                for (int r = 0; r < JRY; ++r) nxt[r] = A[c + row[r] + 2 * sz];
                ym_n = A[c - sy + sz]; yp_n = A[c + row_p + sz];
            }
        } else {
#pragma unroll
            for (int r = 0; r < JRY; ++r) up[r] = A[c + row[r] + sz];
            ym = A[c - sy]; yp = A[c + row_p];
        }
#pragma unroll
        for (int r = 0; r < JRY; ++r) {
            double xm = __shfl_up_sync(0xffffffffu, cen[r], 1), xp = __shfl_down_sync(0xffffffffu, cen[r], 1);
            if (edge_m) xm = A[c + row[r] - 1];
            if (edge_p) xp = A[c + row[r] + 1];
            double s = xm + xp;
            s += (r == 0) ? ym : cen[r - 1];
            s += (r == JRY - 1) ? yp : cen[r + 1];
            s += below[r];
            s += up[r];
            s += HAS_F ? f[c + row[r]] : 0.0;
            if (active && j0 + r <= g.ny) {
                const double val = (1.0 / 6.0) * s;
                B[c + row[r]] = val;
                if (PEER && on_face) jac_peer_store(P, g, i, j0 + r, k, val);
            }
        }
#pragma unroll
        for (int r = 0; r < JRY; ++r) { below[r] = cen[r]; cen[r] = up[r]; }
        c += sz;
    }
}

// ---- the TMA pipeline --------------------------------------------------------------------------------------------------
// Tile = TBX x th cells of one z plane (th <= TBY rows, chosen per block size); a stage holds the tile with its rim:
// (TBX + 2 XH) x (th + 2) doubles, dense.  Work item = (tile, z chunk); item q = chunk * ntiles + tile goes to CTA q mod grid, and
// (th, chunks) are chosen so that the items fill the SMs in whole waves (512 x 512 planes: th = 14 -> 4 x 37 = 148 tiles, one per
// SM).  All CTAs therefore march z in lockstep: a row that two neighbouring tiles both need is fetched from DRAM once and hits
// the L2 for the other (a CTA streams 2.8 MB per plane across the chip; the L2 holds ~20 planes of history -- an earlier
// balanced split that let neighbours drift 20-35 planes apart re-read every shared row from DRAM, profiles/r2c_*).
// A CTA = 128 x NRG consumer threads + one producer warp; a consumer owns TRY consecutive rows of one x column, so the
// largest tile height is TRY * NRG.  Two shapes are built: 4 rows x 4 groups (16 warps, th <= 16) and 2 rows x 7 groups
// (28 warps, th <= 14: twice the warps per scheduler to cover the shared-memory latency of the same work).
constexpr int TBX = 128;
constexpr int TBY_MAX = 16;
// XH = x rim of a stage in cells.  One cell would do for the stencil, but the first column of a box must start on a 16-byte
// boundary in global memory (cp.async.bulk.tensor raised "illegal instruction" on B200 with the box at x index OX - 1 = 15, an
// odd multiple of 8 bytes); with XH = 2 it starts at OX - 2 + 128 t.
constexpr int TMA_XH = 2;
constexpr int TSX = TBX + 2 * TMA_XH;
__host__ __device__ constexpr int tma_stage_bytes(int th) { return ((TSX * (th + 2) * 8 + 127) / 128) * 128; }
constexpr int tma_smem(int stages, int thmax) { return stages * tma_stage_bytes(thmax) + 2 * stages * 8 + 128; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_plane(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

template <bool HAS_F, bool PEER, int TMA_STAGES, int TRY, int NRG, int CTAS>
__global__ void __launch_bounds__(TBX * NRG + 32, CTAS) k_jacobi3d_tma(const __grid_constant__ CUtensorMap mapP,
                                                                 const CUtensorMap *__restrict__ mapG, Geom g,
                                                                 const double *__restrict__ f, double *__restrict__ B,
                                                                 int k_lo, int k_hi, int th, int nchunks, JacPeers P) {
    constexpr int CONSUMERS = TBX * NRG;
    const bool peers_ok = PEER && *P.err == 0;                 // after a failed barrier nothing is stored into a neighbour
    const CUtensorMap *map = mapG ? mapG : &mapP;              // descriptor in global memory, or the kernel parameter
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int stage_bytes = tma_stage_bytes(th);
    uint64_t *full = (uint64_t *)(base + TMA_STAGES * stage_bytes), *empty = full + TMA_STAGES;
    const int tiles_x = (g.nx + TBX - 1) / TBX, tiles_y = (g.ny + th - 1) / th, ntiles = tiles_x * tiles_y;
    const int nz = k_hi - k_lo + 1, H = (nz + nchunks - 1) / nchunks, items = ntiles * nchunks;
    if (threadIdx.x == 0) {
        for (int q = 0; q < TMA_STAGES; ++q) { mbar_init(full + q, 1); mbar_init(empty + q, CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    unsigned n = 0;                                            // running plane counter: stage = n % STAGES
    if (threadIdx.x >= CONSUMERS) {
        // ---- producer: one thread streams planes ka-1 .. kb+1 of every item into the ring
        if (threadIdx.x == CONSUMERS) {
            for (int q = blockIdx.x; q < items; q += gridDim.x) {
                const int t = q % ntiles, ka = k_lo + (q / ntiles) * H, kb = min(k_hi, ka + H - 1);
                const int x0 = (t % tiles_x) * TBX + OX - TMA_XH, y0 = (t / tiles_x) * th;      // cell (i0 - XH, j0 - 1)
                for (int k = ka - 1; k <= kb + 1; ++k, ++n) {
                    const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
                    mbar_wait(empty + st, ph ^ 1u);
                    mbar_expect_tx(full + st, TSX * (th + 2) * 8);
                    tma_load_plane(base + st * stage_bytes, map, full + st, x0, y0, k);
                }
            }
        }
        return;
    }
    // ---- consumers: thread = one x column, TRY consecutive rows of the tile.  Stage row 0 is the rim below the tile, rows
    // 1..th the tile, row th+1 the rim above; a thread row past the tile folds onto the rim above (read, never stored), so the
    // last row of the tile always sees its true upper neighbour
    const int lx = threadIdx.x % TBX, rg = threadIdx.x / TBX;
    int so[TRY];                                               // my cells inside a stage
#pragma unroll
    for (int r = 0; r < TRY; ++r) so[r] = (min(rg * TRY + r, th) + 1) * TSX + lx + TMA_XH;
    const int so_m = min(rg * TRY, th + 1) * TSX + lx + TMA_XH;                     // the row below my strip
    const int so_p = (min(rg * TRY + TRY, th) + 1) * TSX + lx + TMA_XH;             // the row above it
    const bool lane0 = (threadIdx.x & 31) == 0;
    for (int q = blockIdx.x; q < items; q += gridDim.x) {
        const int t = q % ntiles, ka = k_lo + (q / ntiles) * H, kb = min(k_hi, ka + H - 1);
        const int i = 1 + (t % tiles_x) * TBX + lx, j0 = 1 + (t / tiles_x) * th + rg * TRY;
        const bool active = i <= g.nx;
        const int tx = t % tiles_x, ty = t / tiles_x;
        const bool on_face = peers_ok && ((tx == 0) | (tx == tiles_x - 1) | (ty == 0) | (ty == tiles_y - 1) | (ka == 1) | (kb == g.nz));
        long long c = g.idx(0, min(i, g.nx), min(j0, g.ny + 1), ka);
        double below[TRY], cen[TRY], up[TRY];
        {   // plane ka-1: only my own cells, then the stage is free again
            const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
            mbar_wait(full + st, ph);
            const double *S = (const double *)(base + st * stage_bytes);
#pragma unroll
            for (int r = 0; r < TRY; ++r) below[r] = S[so[r]];
            __syncwarp();
            if (lane0) mbar_arrive(empty + st);
            ++n;
        }
        {   // plane ka
            const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
            mbar_wait(full + st, ph);
            const double *S = (const double *)(base + st * stage_bytes);
#pragma unroll
            for (int r = 0; r < TRY; ++r) cen[r] = S[so[r]];
        }
        for (int k = ka; k <= kb; ++k) {
            const unsigned st = n % TMA_STAGES;                                           // holds plane k (already waited for)
            const unsigned su = (n + 1) % TMA_STAGES, pu = ((n + 1) / TMA_STAGES) & 1u;   // plane k+1
            const double *S = (const double *)(base + st * stage_bytes);
            const double *U = (const double *)(base + su * stage_bytes);
            mbar_wait(full + su, pu);
#pragma unroll
            for (int r = 0; r < TRY; ++r) up[r] = U[so[r]];
            const double ym = S[so_m], yp = S[so_p];
#pragma unroll
            for (int r = 0; r < TRY; ++r) {
                double s = S[so[r] - 1] + S[so[r] + 1];
                s += (r == 0) ? ym : cen[r - 1];
                s += (r == TRY - 1) ? yp : cen[r + 1];
                s += below[r];
                s += up[r];
                const long long cr = c + (long long)r * g.sy;
                if (active && rg * TRY + r < th && j0 + r <= g.ny) {
                    s += HAS_F ? f[cr] : 0.0;
                    const double val = (1.0 / 6.0) * s;
                    B[cr] = val;
                    if (PEER && on_face) jac_peer_store(P, g, i, j0 + r, k, val);
                }
            }
#pragma unroll
            for (int r = 0; r < TRY; ++r) { below[r] = cen[r]; cen[r] = up[r]; }
            __syncwarp();
            if (lane0) mbar_arrive(empty + st);                // plane k is no longer needed by this warp
            ++n;
            c += g.sz;
        }
        // plane kb+1 was only read as `up`
        __syncwarp();
        if (lane0) mbar_arrive(empty + (n % TMA_STAGES));
        ++n;
    }
}
// the shapes that are built: index 0 = 4 rows x 4 groups, 10 stages, one CTA per SM; 1 = 2 rows x 7 groups, 10 stages, one CTA per SM;
// 2 = 4 rows x 4 groups, 5 stages, two CTAs per SM
struct TmaShape { int try_, nrg, stages, ctas; };
static const TmaShape TMA_SHAPES[3] = {{4, 4, 10, 1}, {2, 7, 10, 1}, {4, 4, 5, 2}};
template <bool HAS_F, bool PEER>
static cudaError_t launch_jacobi_tma(int shape, int grid, cudaStream_t s, const CUtensorMap &mapP, const CUtensorMap *mapG, const Geom &g,
                                     const double *f, double *B, int k_lo, int k_hi, int th, int nch, const JacPeers &P, bool attr_only) {
#define MGLC_TMA_CASE(IDX, NST, TRY_, NRG_, CTAS_)                                                                                       \
    if (shape == IDX) {                                                                                                                 \
        auto kern = k_jacobi3d_tma<HAS_F, PEER, NST, TRY_, NRG_, CTAS_>;                                                                \
        if (attr_only) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem(NST, TRY_ * NRG_));     \
        kern<<<grid, TBX * NRG_ + 32, tma_smem(NST, TRY_ * NRG_), s>>>(mapP, mapG, g, f, B, k_lo, k_hi, th, nch, P);                    \
        return cudaSuccess;                                                                                                             \
    }
    MGLC_TMA_CASE(0, 10, 4, 4, 1)
    MGLC_TMA_CASE(1, 10, 2, 7, 1)
    MGLC_TMA_CASE(2, 5, 4, 4, 2)
#undef MGLC_TMA_CASE
    return cudaErrorInvalidValue;
}

// planes a CTA marches: short chunks keep the last wave of CTAs small (the sweep of a 512^3 block lasts only
// ~0.4 ms, so a partly filled last wave costs several per cent), long chunks re-read fewer start-up planes
// rows a thread owns (4 or 8): more rows re-read fewer y-neighbour rows from the L2 but need more registers -- measured on
// B200 at 512^3 (profiles/r1k_jacobi_register_blocking_sweep.json): 4 rows 0.371 ms/sweep, 8 rows 0.580 ms (127 registers halve
// the resident warps; the sweep is bound by loads in flight, not by L2 traffic), so 4 stays the default
static int jacobi_jry() {
    static int v = 0;
    if (!v) { v = 4; if (const char *e = getenv("MGLC_JACOBI_JRY")) v = atoi(e) == 8 ? 8 : 4; }
    return v;
}
static int jacobi_kch(int nz) {
    static int v = 0;
    if (!v) { v = 8; if (const char *e = getenv("MGLC_JACOBI_KCH")) v = std::max(1, atoi(e)); }
    return std::min(v, nz);
}

// max |A_p - A| over the interior; non-negative doubles order like their bit patterns
__global__ void __launch_bounds__(256) k_absdiff_max(Geom g, int ndim, const double *__restrict__ A,
                                                     const double *__restrict__ Ap, unsigned long long *out) {
    const long long n = (long long)g.nx * g.ny * g.nz;
    double e = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(q % g.nx), j = 1 + (int)((q / g.nx) % g.ny), k = 1 + (int)(q / ((long long)g.nx * g.ny));
        const long long c = g.idx(0, i, j, k);
        e = fmax(e, fabs(Ap[c] - A[c]));
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = e;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(sm[0]));
}

__global__ void k_fill_range(double *p, long long n, double v) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) p[q] = v;
}

}  // namespace

struct mglc_jacobi {
    int ndim, gn[3], dims[3], nranks;
    std::vector<JacSub *> subs;     // the subdomains this process owns (all of them, or exactly one)
    std::vector<Port> ports;
    mglc_comm *comm;
    int halo_mode;                  // 1 = direct halo stores in the fused step when every mapping came up, 0 = exchange, then sweep
};

// ---- geometry helpers ---------------------------------------------------------------------------------
static inline long long jac_doubles(const JacSub *S) { return S->g.sq; }
static void face_dims(const JacSub *S, int ndim, int face, int &n1, int &n2) {
    const int axis = face >> 1;
    n1 = (axis == 0) ? S->n[1] : S->n[0];
    n2 = (ndim == 2) ? 1 : ((axis == 2) ? S->n[1] : S->n[2]);
}

extern "C" int mglc_dims_create_nd(int nranks, int ndim, int dims[3]) {
    if (nranks < 1 || !dims || (ndim != 2 && ndim != 3)) { set_error("mglc_dims_create_nd: nranks=%d ndim=%d", nranks, ndim); return MGLC_E_INVALID; }
    if (ndim == 3) return mglc_dims_create(nranks, dims);
    int best[3] = {nranks, 1, 1};                          // MPI_Dims_create(np, 2, dims), LAP:47
    for (int a = 1; a <= nranks; ++a)
        if (nranks % a == 0 && nranks / a <= a && a < best[0]) { best[0] = a; best[1] = nranks / a; }
    memcpy(dims, best, sizeof best);
    return MGLC_OK;
}

static int jac_use(JacSub *S) { MGLC_CUDA(cudaSetDevice(S->device)); return MGLC_OK; }
static int jac_wait_direct(mglc_jacobi *h, JacSub *S);
static int jac_drain(mglc_jacobi *h);

// 3-D tiled tensor maps over the padded arrays for k_jacobi3d_tma: dims (px, py, pz) doubles, box (TBX+2, TBY+2, 1), no
// swizzle (the stage is read row-wise by consecutive lanes: conflict-free as it is), out-of-bounds = 0 (tiles overhanging
// the block; those cells are never stored).  cuTensorMapEncodeTiled is looked up in the driver at run time.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}
static int sm_count(int device);
// which of TMA_SHAPES a handle uses (fixed when it is created: the tile height is part of the tensor map)
static int jacobi_tma_shape_index() {
    const char *e = getenv("MGLC_JACOBI_TMA_SHAPE");
    const int v = e ? atoi(e) : 0;
    return v >= 0 && v <= 2 ? v : 0;
}
// tile height th and number of z chunks: the work items (tiles x chunks) should fill the CTAs in whole waves, with as little rim
// ((th + 2) / th rows loaded per row computed) and as few chunk restarts (2 extra planes per chunk) as that allows
static void jac_tma_shape(JacSub *S) {
    S->tma_shape = jacobi_tma_shape_index();
    const TmaShape &sh = TMA_SHAPES[S->tma_shape];
    const int thmax = sh.try_ * sh.nrg;
    const int G = sm_count(S->device) * sh.ctas;
    const int tiles_x = (S->n[0] + TBX - 1) / TBX;
    double best = -1.0;
    for (int th = thmax; th >= 4; --th)
        for (int nch = 1; nch <= 32; ++nch) {
            const int H = (S->n[2] + nch - 1) / nch;
            if (nch > 1 && H < 16) break;
            const long long items = (long long)tiles_x * ((S->n[1] + th - 1) / th) * nch;
            const long long waves = (items + G - 1) / G;
            const double score = (double)items / (double)(waves * G) * th / (th + 2.0) * H / (H + 2.0);
            if (score > best * 1.0001) { best = score; S->tma_th = th; S->tma_chunks = nch; }
        }
    if (const char *e = getenv("MGLC_JACOBI_TMA_TH")) S->tma_th = std::max(1, std::min(thmax, atoi(e)));
    if (const char *e = getenv("MGLC_JACOBI_TMA_CHUNKS")) S->tma_chunks = std::max(1, std::min(S->n[2], atoi(e)));
}
static void jac_make_tmaps(JacSub *S) {
    S->tma_ok = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return;
    jac_tma_shape(S);
    const cuuint64_t dims[3] = {(cuuint64_t)S->g.px, (cuuint64_t)S->g.py, (cuuint64_t)S->g.pz};
    const cuuint64_t strides[2] = {(cuuint64_t)S->g.sy * 8, (cuuint64_t)S->g.sz * 8};
    const cuuint32_t box[3] = {TSX, (cuuint32_t)(S->tma_th + 2), 1}, estr[3] = {1, 1, 1};
    // no L2 promotion: with 256-byte promotion the rows of a box (1056 B, starting 112 B into a line) pull in more DRAM sectors
    // than the SMs ever request (ncu, 512^3: 1.44 GB read from DRAM for 1.37 GB delivered; profiles/r2c_*)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (const char *e = getenv("MGLC_JACOBI_TMA_L2PROMO")) {
        const int v = atoi(e);
        promo = v >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : v >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : promo;
    }
    for (int b = 0; b < 2; ++b)
        if (enc(&S->tmap[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, S->A[b], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return;
    {
        const JacPeers none{};
        const Geom &g = S->g;
        if (launch_jacobi_tma<false, false>(S->tma_shape, 0, nullptr, S->tmap[0], nullptr, g, nullptr, nullptr, 0, 0, 0, 0, none, true) != cudaSuccess ||
            launch_jacobi_tma<true, false>(S->tma_shape, 0, nullptr, S->tmap[0], nullptr, g, nullptr, nullptr, 0, 0, 0, 0, none, true) != cudaSuccess ||
            launch_jacobi_tma<false, true>(S->tma_shape, 0, nullptr, S->tmap[0], nullptr, g, nullptr, nullptr, 0, 0, 0, 0, none, true) != cudaSuccess ||
            launch_jacobi_tma<true, true>(S->tma_shape, 0, nullptr, S->tmap[0], nullptr, g, nullptr, nullptr, 0, 0, 0, 0, none, true) != cudaSuccess) { (void)cudaGetLastError(); return; }
    }
    if (cudaMalloc((void **)&S->tmap_dev, 2 * sizeof(CUtensorMap)) != cudaSuccess ||
        cudaMemcpy(S->tmap_dev, S->tmap, 2 * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) { (void)cudaGetLastError(); return; }
    S->tma_ok = 1;
}
// MGLC_JACOBI_KERNEL=tma selects the TMA pipeline; default: the register-blocked LDG kernel (k_jacobi3d), which is the faster of
// the two on B200 at 512^3 (0.367 ms against 0.424 ms per sweep, profiles/r2e_*)
// (read at every sweep, so a test or a tuning sweep can flip it inside one process; MGLC_JACOBI_TMA_SHAPE / _TH / _CHUNKS are read
// when the handle is created, because the tile height is part of the tensor map)
static bool jacobi_use_tma() {
    const char *e = getenv("MGLC_JACOBI_KERNEL");
    return e && !strcmp(e, "tma");
}
static int sm_count(int device) {
    static int n[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (!n[device]) { cudaDeviceGetAttribute(&n[device], cudaDevAttrMultiProcessorCount, device); if (n[device] <= 0) n[device] = 148; }
    return n[device];
}

static void jac_free_sub(JacSub *S) {
    if (!S) return;
    cudaSetDevice(S->device);
    if (S->s) cudaStreamSynchronize(S->s);
    double *bufs[] = {S->A[0], S->A[1], S->f, S->A_p, S->scratch};
    for (double *p : bufs) cudaFree(p);
    cudaFree(S->tmap_dev);
    if (S->ipc_opened) { for (void *q : *S->ipc_opened) cudaIpcCloseMemHandle(q); delete S->ipc_opened; }
    cudaFree(S->flags); cudaFree(S->d_err);
    for (cudaEvent_t e : S->ev_done) if (e) cudaEventDestroy(e);
    for (Msg &M : S->msgs) { cudaFree(M.sbuf); cudaFree(M.rbuf); }
    cudaEvent_t evs[] = {S->ev_packed, S->ev_copied, S->ev_t0, S->ev_t1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (S->s) cudaStreamDestroy(S->s);
    delete S;
}

extern "C" int mglc_jacobi_destroy(mglc_jacobi *h) {
    if (!h) return MGLC_OK;
    // neighbours store into each other's ghost layers: quiesce everybody before anything is freed
    if (h->comm) for (JacSub *S : h->subs) { cudaSetDevice(S->device); if (S->direct_valid && S->s) { launch_halo_wait(S->sync, S->epoch * 2 + (unsigned long long)S->parity_sent, S->d_err, S->s); S->direct_valid = 0; } }
    for (JacSub *S : h->subs) { cudaSetDevice(S->device); cudaDeviceSynchronize(); }
    for (JacSub *S : h->subs) jac_free_sub(S);
    delete h;
    return MGLC_OK;
}

static int jac_make_sub(mglc_jacobi *h, int rank, int device, JacSub **out) {
    JacSub *S = new JacSub();
    memset(S, 0, sizeof *S);
    S->device = device;
    S->coords[2] = rank % h->dims[2];
    S->coords[1] = (rank / h->dims[2]) % h->dims[1];
    S->coords[0] = rank / (h->dims[2] * h->dims[1]);
    for (int d = 0; d < 3; ++d) {
        if (h->gn[d] < h->dims[d]) { set_error("mglc_jacobi_create: fewer cells than ranks along dim %d", d); delete S; return MGLC_E_INVALID; }
        mglc_decompose_1d(h->gn[d], S->coords[d], h->dims[d], &S->n[d], &S->start[d]);
    }
    for (int d = 0; d < 3; ++d) {
        int p[3] = {S->coords[0], S->coords[1], S->coords[2]}, m[3] = {S->coords[0], S->coords[1], S->coords[2]};
        p[d] += 1; m[d] -= 1;
        mglc_cart_rank(h->dims, p, &S->nbr[2 * d]);
        mglc_cart_rank(h->dims, m, &S->nbr[2 * d + 1]);
    }
    S->g = make_geom(S->n[0], S->n[1], S->n[2]);
    auto fail = [&](int rc) { jac_free_sub(S); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return fail(MGLC_E_CUDA); }
    if (cudaStreamCreateWithFlags(&S->s, cudaStreamNonBlocking) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaEventCreateWithFlags(&S->ev_packed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&S->ev_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&S->ev_t0) != cudaSuccess || cudaEventCreate(&S->ev_t1) != cudaSuccess) return fail(MGLC_E_CUDA);
    const size_t bytes = (size_t)jac_doubles(S) * sizeof(double);
    double **bufs[] = {&S->A[0], &S->A[1], &S->A_p};
    for (double **b : bufs) {
        if (cudaMalloc((void **)b, bytes) != cudaSuccess) { (void)cudaGetLastError(); set_error("mglc_jacobi_create: out of device memory"); return fail(MGLC_E_NOMEM); }
        cudaMemsetAsync(*b, 0, bytes, S->s);
    }
    if (cudaMalloc((void **)&S->scratch, 64) != cudaSuccess) return fail(MGLC_E_NOMEM);
    if (cudaMalloc((void **)&S->flags, 32 * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc((void **)&S->d_err, sizeof(int)) != cudaSuccess) return fail(MGLC_E_NOMEM);
    cudaMemsetAsync(S->flags, 0, 32 * sizeof(unsigned long long), S->s);
    cudaMemsetAsync(S->d_err, 0, sizeof(int), S->s);
    for (cudaEvent_t &e : S->ev_done) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(MGLC_E_CUDA);
    for (int face = 0; face < 2 * h->ndim; ++face) {
        Msg &M = S->msgs[face];
        int n1, n2;
        face_dims(S, h->ndim, face, n1, n2);
        M.dir = face;
        M.send_to = S->nbr[face];
        M.recv_from = S->nbr[face ^ 1];
        M.send_count = M.send_to >= 0 ? (long long)n1 * n2 : 0;
        M.recv_count = M.recv_from >= 0 ? (long long)n1 * n2 : 0;
        if (M.send_count && cudaMalloc((void **)&M.sbuf, M.send_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
        if (M.recv_count && cudaMalloc((void **)&M.rbuf, M.recv_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
    }
    if (cudaStreamSynchronize(S->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (h->ndim == 3) jac_make_tmaps(S);
    *out = S;
    return MGLC_OK;
}

static int jac_new(mglc_jacobi **out, int ndim, const int gn[3], const int dims_or_zero[3], int nranks) {
    if (!out || !gn || (ndim != 2 && ndim != 3) || nranks < 1) { set_error("mglc_jacobi_create: bad arguments"); return MGLC_E_INVALID; }
    for (int d = 0; d < ndim; ++d) if (gn[d] < 1) { set_error("mglc_jacobi_create: gn[%d]=%d", d, gn[d]); return MGLC_E_INVALID; }
    MGLC_TRY(require_gpu());
    mglc_jacobi *h = new mglc_jacobi();
    h->ndim = ndim; h->nranks = nranks; h->comm = nullptr;
    h->gn[0] = gn[0]; h->gn[1] = gn[1]; h->gn[2] = ndim == 3 ? gn[2] : 1;
    if (dims_or_zero && dims_or_zero[0] > 0) memcpy(h->dims, dims_or_zero, sizeof h->dims);
    else mglc_dims_create_nd(nranks, ndim, h->dims);
    if (h->dims[0] * h->dims[1] * h->dims[2] != nranks || (ndim == 2 && h->dims[2] != 1)) {
        set_error("mglc_jacobi_create: dims %dx%dx%d do not fit %d ranks in %d-D", h->dims[0], h->dims[1], h->dims[2], nranks, ndim);
        delete h;
        return MGLC_E_INVALID;
    }
    *out = h;
    return MGLC_OK;
}

// ---- direct halo stores: wiring ----------------------------------------------------------------------------------------
static void jac_install_peer(mglc_jacobi *h, JacSub *S, int face, double *A0, double *A1, unsigned long long *flags, const int ln[3]) {
    S->peerA[face][0] = A0; S->peerA[face][1] = A1;
    S->peer_g[face] = make_geom(ln[0], ln[1], ln[2]);
    S->sync.mask |= 1u << face;
    S->sync.signal[face] = flags + (face ^ 1);          // the neighbour sees me across the opposite face
    S->sync.wait[face] = S->flags + face;
}
// one process per GPU: exchange CUDA IPC handles of both arrays and the barrier words through the communicator, map the
// face neighbours' allocations (peer access over NVLink) and agree collectively whether the path is usable
struct JacIpcRecord { cudaIpcMemHandle_t A[2], flags; int ln[3]; };
static int jac_setup_direct_ipc(mglc_jacobi *h) {
    JacSub *S = h->subs[0];
    if (!h->comm || h->nranks < 2 || getenv("MGLC_NO_DIRECT")) return MGLC_OK;
    MGLC_TRY(jac_use(S));
    const int P = h->nranks;
    JacIpcRecord mine;
    memset(&mine, 0, sizeof mine);
    int ok = 1;
    const int my_rank = (S->coords[0] * h->dims[1] + S->coords[1]) * h->dims[2] + S->coords[2];
    const unsigned long long magic = 0x6a61636f00000000ull + (unsigned long long)my_rank;
    ok &= cudaMemcpy(S->flags + 31, &magic, sizeof magic, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int b = 0; b < 2; ++b) ok &= cudaIpcGetMemHandle(&mine.A[b], S->A[b]) == cudaSuccess;
    ok &= cudaIpcGetMemHandle(&mine.flags, S->flags) == cudaSuccess;
    (void)cudaGetLastError();
    for (int q = 0; q < 3; ++q) mine.ln[q] = S->n[q];
    char *dev_all = nullptr;
    MGLC_CUDA(cudaMalloc((void **)&dev_all, (size_t)(P + 1) * sizeof(JacIpcRecord)));
    MGLC_CUDA(cudaMemcpy(dev_all + (size_t)P * sizeof(JacIpcRecord), &mine, sizeof mine, cudaMemcpyHostToDevice));
    MGLC_NCCL(ncclAllGather(dev_all + (size_t)P * sizeof(JacIpcRecord), dev_all, sizeof(JacIpcRecord), ncclChar, h->comm->nccl, S->s));
    std::vector<JacIpcRecord> all(P);
    MGLC_CUDA(cudaMemcpyAsync(all.data(), dev_all, (size_t)P * sizeof(JacIpcRecord), cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    cudaFree(dev_all);
    S->ipc_opened = new std::vector<void *>();
    auto open = [&](const cudaIpcMemHandle_t &hd, void **out) {
        if (cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); *out = nullptr; return 0; }
        S->ipc_opened->push_back(*out);
        return 1;
    };
    memset(&S->sync, 0, sizeof S->sync);
    for (int f = 0; f < 2 * h->ndim && ok; ++f) {
        const int r = S->nbr[f];
        if (r < 0) continue;
        // with two ranks along an axis the +face and the -face neighbour are different ranks only if dims > 2; the same rank
        // may be opened twice (once per face), which cudaIpcOpenMemHandle does not allow: reuse the first mapping
        int prev = -1;
        for (int q = 0; q < f; ++q) if (S->nbr[q] == r) prev = q;
        double *A0 = nullptr, *A1 = nullptr;
        unsigned long long *fl = nullptr;
        if (prev >= 0) { A0 = S->peerA[prev][0]; A1 = S->peerA[prev][1]; fl = S->sync.signal[prev] - (prev ^ 1); }
        else {
            ok &= open(all[r].A[0], (void **)&A0);
            if (ok) ok &= open(all[r].A[1], (void **)&A1);
            if (ok) ok &= open(all[r].flags, (void **)&fl);
            if (ok) {           // the mapping must start at the neighbour's own pointer, not at some enclosing block
                unsigned long long seen = 0;
                ok &= cudaMemcpy(&seen, fl + 31, sizeof seen, cudaMemcpyDeviceToHost) == cudaSuccess && seen == 0x6a61636f00000000ull + (unsigned long long)r;
                (void)cudaGetLastError();
            }
        }
        if (ok) jac_install_peer(h, S, f, A0, A1, fl, all[r].ln);
    }
    int *dev_ok = nullptr;
    MGLC_CUDA(cudaMalloc((void **)&dev_ok, sizeof(int)));
    MGLC_CUDA(cudaMemcpy(dev_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
    MGLC_NCCL(ncclAllReduce(dev_ok, dev_ok, 1, ncclInt, ncclMin, h->comm->nccl, S->s));
    MGLC_CUDA(cudaMemcpyAsync(&ok, dev_ok, sizeof ok, cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    cudaFree(dev_ok);
    S->direct = ok ? 1 : 0;
    return MGLC_OK;
}
// P subdomains in one process: plain device pointers (same device, or peer access enabled), ordering through events
static void jac_setup_direct_local(mglc_jacobi *h) {
    bool reachable = h->nranks > 1 && !getenv("MGLC_NO_DIRECT");
    for (JacSub *a : h->subs)
        for (JacSub *b : h->subs)
            if (reachable && a->device != b->device) { int can = 0; cudaDeviceCanAccessPeer(&can, a->device, b->device); reachable = can != 0; }
    if (!reachable) return;
    for (JacSub *S : h->subs) {
        memset(&S->sync, 0, sizeof S->sync);
        for (int f = 0; f < 2 * h->ndim; ++f) {
            if (S->nbr[f] < 0) continue;
            JacSub *N = h->subs[S->nbr[f]];
            jac_install_peer(h, S, f, N->A[0], N->A[1], N->flags, N->n);
        }
        S->direct = 1;
    }
}

extern "C" int mglc_jacobi_create(mglc_jacobi **out, int ndim, const int gn[3], const int dims_or_zero[3], int nranks,
                                  int rank, int device, mglc_comm *comm_or_null) {
    if (nranks > 1 && !comm_or_null) { set_error("mglc_jacobi_create: %d ranks need a communicator (or use mglc_jacobi_create_local)", nranks); return MGLC_E_INVALID; }
    if (rank < 0 || rank >= nranks) { set_error("mglc_jacobi_create: rank=%d of %d", rank, nranks); return MGLC_E_INVALID; }
    mglc_jacobi *h = nullptr;
    MGLC_TRY(jac_new(&h, ndim, gn, dims_or_zero, nranks));
    h->comm = comm_or_null;
    JacSub *S = nullptr;
    int rc = jac_make_sub(h, rank, device, &S);
    if (rc) { delete h; return rc; }
    h->subs.push_back(S);
    h->halo_mode = 1;
    if (nranks > 1) {
        rc = jac_setup_direct_ipc(h);                    // collective over the communicator; failure to map = NCCL exchange
        if (rc) { mglc_jacobi_destroy(h); return rc; }
    }
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_jacobi_create_local(mglc_jacobi **out, int ndim, const int gn[3], const int dims_or_zero[3],
                                        int nranks, const int *devices_or_null) {
    mglc_jacobi *h = nullptr;
    MGLC_TRY(jac_new(&h, ndim, gn, dims_or_zero, nranks));
    for (int r = 0; r < nranks; ++r) {
        JacSub *S = nullptr;
        int rc = jac_make_sub(h, r, devices_or_null ? devices_or_null[r] : 0, &S);
        if (rc) { mglc_jacobi_destroy(h); return rc; }
        h->subs.push_back(S);
    }
    for (JacSub *a : h->subs)
        for (JacSub *b : h->subs)
            if (a->device != b->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->device, b->device);
                if (can) { cudaSetDevice(a->device); cudaDeviceEnablePeerAccess(b->device, 0); (void)cudaGetLastError(); }
            }
    for (JacSub *S : h->subs) h->ports.push_back(Port{S->device, S->s, S->ev_packed, S->ev_copied, S->msgs, 2 * ndim});
    h->halo_mode = 1;
    jac_setup_direct_local(h);
    *out = h;
    return MGLC_OK;
}

static int jac_sub(mglc_jacobi *h, int r, JacSub **S) {
    if (!h || r < 0 || r >= (int)h->subs.size()) { set_error("mglc_jacobi: bad handle or local index %d", r); return MGLC_E_INVALID; }
    *S = h->subs[r];
    return jac_use(*S);
}

extern "C" int mglc_jacobi_nlocal(mglc_jacobi *h, int *n) {
    if (!h || !n) return MGLC_E_INVALID;
    *n = (int)h->subs.size();
    return MGLC_OK;
}
extern "C" int mglc_jacobi_info(mglc_jacobi *h, int r, int dims[3], int ln[3], int start[3], int coords[3], int nbr[6]) {
    if (!h || r < 0 || r >= (int)h->subs.size()) return MGLC_E_INVALID;
    JacSub *S = h->subs[r];
    if (dims) memcpy(dims, h->dims, 12);
    if (ln) memcpy(ln, S->n, 12);
    if (start) memcpy(start, S->start, 12);
    if (coords) memcpy(coords, S->coords, 12);
    if (nbr) memcpy(nbr, S->nbr, 24);
    return MGLC_OK;
}

// host (0:nx+1, 0:ny+1 [, 0:nz+1]) column-major <-> padded device rows; one strided copy
static int jac_copy(mglc_jacobi *h, JacSub *S, double *host, double *dev, bool to_device) {
    if (!host) return MGLC_OK;
    const size_t w = (size_t)(S->n[0] + 2) * sizeof(double);
    const size_t rows = (h->ndim == 3) ? (size_t)S->g.py * S->g.pz : (size_t)S->g.py;
    double *d0 = dev + (OX - 1) + (h->ndim == 3 ? 0 : S->g.sz);
    if (to_device) MGLC_CUDA(cudaMemcpy2DAsync(d0, (size_t)S->g.px * sizeof(double), host, w, w, rows, cudaMemcpyHostToDevice, S->s));
    else MGLC_CUDA(cudaMemcpy2DAsync(host, w, d0, (size_t)S->g.px * sizeof(double), w, rows, cudaMemcpyDeviceToHost, S->s));
    return MGLC_OK;
}

extern "C" int mglc_jacobi_upload(mglc_jacobi *h, int r, const double *A, const double *A_new, const double *f) {
    JacSub *S;
    MGLC_TRY(jac_sub(h, r, &S));
    MGLC_TRY(jac_wait_direct(h, S));
    MGLC_TRY(jac_copy(h, S, const_cast<double *>(A), S->A[S->cur], true));
    MGLC_TRY(jac_copy(h, S, const_cast<double *>(A_new), S->A[S->cur ^ 1], true));
    if (f) {
        if (!S->f) {
            MGLC_CUDA(cudaMalloc((void **)&S->f, (size_t)jac_doubles(S) * sizeof(double)));
            MGLC_CUDA(cudaMemsetAsync(S->f, 0, (size_t)jac_doubles(S) * sizeof(double), S->s));
        }
        MGLC_TRY(jac_copy(h, S, const_cast<double *>(f), S->f, true));
    }
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
extern "C" int mglc_jacobi_download(mglc_jacobi *h, int r, double *A, double *A_new) {
    JacSub *S;
    MGLC_TRY(jac_sub(h, r, &S));
    MGLC_TRY(jac_wait_direct(h, S));
    MGLC_TRY(jac_copy(h, S, A, S->A[S->cur], false));
    MGLC_TRY(jac_copy(h, S, A_new, S->A[S->cur ^ 1], false));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}

// init(), LAP:144-166: everything 0, the top ghost layer of the top ranks 1 (rims included), A_p = A
extern "C" int mglc_jacobi_init(mglc_jacobi *h) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(jac_drain(h));
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        const size_t bytes = (size_t)jac_doubles(S) * sizeof(double);
        S->cur = 0;
        for (int b = 0; b < 2; ++b) MGLC_CUDA(cudaMemsetAsync(S->A[b], 0, bytes, S->s));
        if (S->f) MGLC_CUDA(cudaMemsetAsync(S->f, 0, bytes, S->s));
        const int top = h->ndim - 1;
        if (S->coords[top] == h->dims[top] - 1) {
            for (int b = 0; b < 2; ++b) {
                if (h->ndim == 2) {       // row j = ny+1 of plane k = 1, i = 0..nx+1
                    double *row = S->A[b] + S->g.idx(0, 0, S->n[1] + 1, 1);
                    k_fill_range<<<(S->n[0] + 2 + 255) / 256, 256, 0, S->s>>>(row, S->n[0] + 2, 1.0);
                } else {                  // plane k = nz+1 (padding columns included: never read)
                    double *plane = S->A[b] + S->g.sz * (S->n[2] + 1);
                    k_fill_range<<<592, 256, 0, S->s>>>(plane, S->g.sz, 1.0);
                }
                S->launches += 1;
            }
        }
        MGLC_CUDA(cudaMemcpyAsync(S->A_p, S->A[0], bytes, cudaMemcpyDeviceToDevice, S->s));
    }
    return MGLC_OK;
}

static int jac_pack(mglc_jacobi *h, JacSub *S, cudaStream_t s) {
    for (int face = 0; face < 2 * h->ndim; ++face) {
        Msg &M = S->msgs[face];
        if (!M.send_count) continue;
        int n1, n2;
        face_dims(S, h->ndim, face, n1, n2);
        k_face_pack<<<(unsigned)((M.send_count + 255) / 256), 256, 0, s>>>(S->g, S->A[S->cur], face, n1, n2, M.sbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}
static int jac_unpack(mglc_jacobi *h, JacSub *S, cudaStream_t s) {
    for (int face = 0; face < 2 * h->ndim; ++face) {
        Msg &M = S->msgs[face];
        if (!M.recv_count) continue;
        int n1, n2;
        face_dims(S, h->ndim, face, n1, n2);
        k_face_unpack<<<(unsigned)((M.recv_count + 255) / 256), 256, 0, s>>>(S->g, S->A[S->cur], face, n1, n2, M.rbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}

// exchange_message(A), LAP:223-254
static int jac_exchange(mglc_jacobi *h) {
    if (h->nranks == 1) return MGLC_OK;
    MGLC_TRY(jac_drain(h));
    if (h->comm) {
        JacSub *S = h->subs[0];
        MGLC_TRY(jac_use(S));
        MGLC_TRY(jac_pack(h, S, S->s));
        MGLC_TRY(halo_nccl_sendrecv(S->msgs, 2 * h->ndim, h->comm, S->s));
        MGLC_TRY(jac_unpack(h, S, S->s));
        return MGLC_OK;
    }
    return halo_local_exchange(
        h->ports, [&](int r, cudaStream_t s) { return jac_pack(h, h->subs[r], s); },
        [&](int r, cudaStream_t s) { return jac_unpack(h, h->subs[r], s); });
}

// ---- direct halo stores: the barrier halves and the peer table of one launch ----------------------------------------------
static JacSub *jac_local_sub(mglc_jacobi *h, int rank) { return h->comm ? nullptr : h->subs[rank]; }
// every neighbour has finished the sweep of my current epoch: its stores into my ghost layers are complete and it no
// longer reads the ghost layers my next sweep will overwrite
static int jac_wait_direct(mglc_jacobi *h, JacSub *S) {
    if (!S->direct_valid) return MGLC_OK;
    if (h->comm) S->launches += launch_halo_wait(S->sync, S->epoch * 2 + (unsigned long long)S->parity_sent, S->d_err, S->s);
    else
        for (int f = 0; f < 2 * h->ndim; ++f)
            if (S->nbr[f] >= 0) MGLC_CUDA(cudaStreamWaitEvent(S->s, jac_local_sub(h, S->nbr[f])->ev_done[S->epoch & 1], 0));
    S->direct_valid = 0;
    return MGLC_OK;
}
// anything that reads or rewrites the arrays outside the fused step first lets the neighbours' stores land
static int jac_drain(mglc_jacobi *h) {
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_TRY(jac_wait_direct(h, S)); }
    return MGLC_OK;
}
static JacPeers jac_peers(mglc_jacobi *h, JacSub *S, bool on) {
    JacPeers P{};
    if (!on || !S->direct) return P;
    for (int f = 0; f < 2 * h->ndim; ++f) {
        if (S->nbr[f] < 0) continue;
        P.mask |= 1u << f;
        P.B[f] = S->peerA[f][S->cur ^ 1];                  // the neighbour's A_new: same ping-pong index as mine
        P.sy[f] = S->peer_g[f].sy; P.sz[f] = S->peer_g[f].sz;
        P.n[f] = f < 2 ? S->peer_g[f].nx : f < 4 ? S->peer_g[f].ny : S->peer_g[f].nz;
    }
    P.err = S->d_err;
    return P;
}

// jacobi(A, A_new), LAP:170-182, then the roles swap (LAP:97-103 ping-pongs the two arrays); peers: also store the
// boundary values into the neighbours' ghost layers and raise the barrier
static int jac_sweep(mglc_jacobi *h, bool peers = false) {
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        const double *A = S->A[S->cur];
        double *B = S->A[S->cur ^ 1];
        const JacPeers P = jac_peers(h, S, peers);
        // default: the plain sweep, then k_jac_push; MGLC_JACOBI_PEER_IN_KERNEL=1 stores from inside the sweep instead
        static const bool in_kernel = getenv("MGLC_JACOBI_PEER_IN_KERNEL") != nullptr;
        const bool peer = P.mask != 0 && in_kernel;
        if (h->ndim == 2) {
            const dim3 grid((S->n[0] + 127) / 128, S->n[1]);
            if (peer) { if (S->f) k_jacobi2d<true, true><<<grid, 128, 0, S->s>>>(S->g, A, S->f, B, P); else k_jacobi2d<false, true><<<grid, 128, 0, S->s>>>(S->g, A, nullptr, B, P); }
            else { if (S->f) k_jacobi2d<true, false><<<grid, 128, 0, S->s>>>(S->g, A, S->f, B, P); else k_jacobi2d<false, false><<<grid, 128, 0, S->s>>>(S->g, A, nullptr, B, P); }
        } else if (S->tma_ok && jacobi_use_tma()) {
            const int tiles = ((S->n[0] + TBX - 1) / TBX) * ((S->n[1] + S->tma_th - 1) / S->tma_th);
            const int grid = (int)std::min<long long>((long long)sm_count(S->device) * TMA_SHAPES[S->tma_shape].ctas, (long long)tiles * S->tma_chunks);
            const char *where = getenv("MGLC_JACOBI_TMAP");
            const CUtensorMap *gmap = (where && !strcmp(where, "global")) ? S->tmap_dev + S->cur : nullptr;
            const CUtensorMap &mp = S->tmap[S->cur];
            if (peer) { if (S->f) launch_jacobi_tma<true, true>(S->tma_shape, grid, S->s, mp, gmap, S->g, S->f, B, 1, S->n[2], S->tma_th, S->tma_chunks, P, false);
                        else launch_jacobi_tma<false, true>(S->tma_shape, grid, S->s, mp, gmap, S->g, nullptr, B, 1, S->n[2], S->tma_th, S->tma_chunks, P, false); }
            else { if (S->f) launch_jacobi_tma<true, false>(S->tma_shape, grid, S->s, mp, gmap, S->g, S->f, B, 1, S->n[2], S->tma_th, S->tma_chunks, P, false);
                   else launch_jacobi_tma<false, false>(S->tma_shape, grid, S->s, mp, gmap, S->g, nullptr, B, 1, S->n[2], S->tma_th, S->tma_chunks, P, false); }
        } else {
            const int kch = jacobi_kch(S->n[2]);
            const int JRY = jacobi_jry();
            const dim3 grid((S->n[0] + JTX - 1) / JTX, (S->n[1] + JRY - 1) / JRY, (S->n[2] + kch - 1) / kch);
            const dim3 block(JTX);
            const char *pfe = getenv("MGLC_JACOBI_PF");
            const bool pf = pfe ? atoi(pfe) != 0 : false;       // measured on B200 at 512^3: 0.408 ms with the prefetch, 0.367 ms without (7 instead of 8 CTAs per SM)
#define MGLC_JAC_REG(HASF, ROWS, PEER_, FPTR) do { if (pf) k_jacobi3d<HASF, ROWS, PEER_, true><<<grid, block, 0, S->s>>>(S->g, A, FPTR, B, 1, S->n[2], kch, P); \
                                                   else k_jacobi3d<HASF, ROWS, PEER_, false><<<grid, block, 0, S->s>>>(S->g, A, FPTR, B, 1, S->n[2], kch, P); } while (0)
            if (JRY == 8) {
                if (peer) { if (S->f) MGLC_JAC_REG(true, 8, true, S->f); else MGLC_JAC_REG(false, 8, true, nullptr); }
                else { if (S->f) MGLC_JAC_REG(true, 8, false, S->f); else MGLC_JAC_REG(false, 8, false, nullptr); }
            } else {
                if (peer) { if (S->f) MGLC_JAC_REG(true, 4, true, S->f); else MGLC_JAC_REG(false, 4, true, nullptr); }
                else { if (S->f) MGLC_JAC_REG(true, 4, false, S->f); else MGLC_JAC_REG(false, 4, false, nullptr); }
            }
#undef MGLC_JAC_REG
        }
        S->launches += 1;
        if (P.mask && !peer) {
            // a corner/edge cell belongs to two or three faces: jac_peer_store sends it to each of them, as the sweep would
            const long long cells = (long long)std::max(S->n[0], S->n[1]) * (h->ndim == 2 ? 1 : std::max(S->n[1], S->n[2]));
            k_jac_push<<<dim3((unsigned)std::min<long long>(256, (cells + 255) / 256), 2 * h->ndim), 256, 0, S->s>>>(S->g, h->ndim, B, P);
            S->launches += 1;
        }
        if (P.mask) {
            const int parity = S->cur ^ 1;                         // of the array just written
            S->epoch += 1;
            S->parity_sent = parity;
            if (h->comm) S->launches += launch_halo_signal(S->sync, S->epoch * 2 + (unsigned long long)parity, S->s);
            else MGLC_CUDA(cudaEventRecord(S->ev_done[S->epoch & 1], S->s));
            S->direct_valid = 1;
        }
        S->cur ^= 1;
    }
    return MGLC_OK;
}

extern "C" int mglc_jacobi_exchange(mglc_jacobi *h) { if (!h) return MGLC_E_INVALID; return jac_exchange(h); }
extern "C" int mglc_jacobi_sweep(mglc_jacobi *h) { if (!h) return MGLC_E_INVALID; MGLC_TRY(jac_drain(h)); return jac_sweep(h); }

static int jac_step_impl(mglc_jacobi *h, int nits) {
    if (nits < 0) { set_error("mglc_jacobi_step: nits=%d", nits); return MGLC_E_INVALID; }
    bool direct = h->nranks > 1 && h->halo_mode == 1;
    for (JacSub *S : h->subs) direct = direct && S->direct;
    for (int it = 0; it < nits; ++it) {
        bool all_valid = direct;
        for (JacSub *S : h->subs) all_valid = all_valid && S->direct_valid;
        if (all_valid) MGLC_TRY(jac_drain(h));                     // the neighbours' last sweeps stored this iteration's ghost layers
        else MGLC_TRY(jac_exchange(h));                            // exchange_message(A), LAP:223-254 (drains first)
        MGLC_TRY(jac_sweep(h, direct));
    }
    return MGLC_OK;
}
extern "C" int mglc_jacobi_step(mglc_jacobi *h, int nits) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(jac_step_impl(h, nits));
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
extern "C" int mglc_jacobi_step_timed(mglc_jacobi *h, int nits, float *ms) {
    if (!h || !ms) return MGLC_E_INVALID;
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t0, S->s)); }
    MGLC_TRY(jac_step_impl(h, nits));
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t1, S->s)); }
    float worst = 0.f;
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        MGLC_CUDA(cudaEventSynchronize(S->ev_t1));
        MGLC_CUDA(cudaGetLastError());
        float t = 0.f;
        MGLC_CUDA(cudaEventElapsedTime(&t, S->ev_t0, S->ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MGLC_OK;
}

// check_diff + MPI_Allreduce(MAX), LAP:105-107,185-204
extern "C" int mglc_jacobi_check_diff(mglc_jacobi *h, double *error_max) {
    if (!h || !error_max) return MGLC_E_INVALID;
    MGLC_TRY(jac_drain(h));
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        MGLC_CUDA(cudaMemsetAsync(S->scratch, 0, 8, S->s));
        k_absdiff_max<<<592, 256, 0, S->s>>>(S->g, h->ndim, S->A[S->cur], S->A_p, (unsigned long long *)S->scratch);
        S->launches += 1;
        MGLC_CUDA(cudaMemcpyAsync(S->A_p, S->A[S->cur], (size_t)jac_doubles(S) * sizeof(double), cudaMemcpyDeviceToDevice, S->s));
    }
    double worst = 0.0;
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        if (h->comm && h->nranks > 1)
            MGLC_NCCL(ncclAllReduce(S->scratch, S->scratch, 1, ncclDouble, ncclMax, h->comm->nccl, S->s));
        double e = 0.0;
        MGLC_CUDA(cudaMemcpyAsync(&e, S->scratch, 8, cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        worst = std::max(worst, e);
    }
    *error_max = worst;
    return MGLC_OK;
}

extern "C" int mglc_jacobi_launch_count(mglc_jacobi *h, long long *n) {
    if (!h || !n) return MGLC_E_INVALID;
    long long t = 0;
    for (JacSub *S : h->subs) t += S->launches;
    *n = t;
    return MGLC_OK;
}
extern "C" int mglc_jacobi_sync(mglc_jacobi *h) {
    if (!h) return MGLC_E_INVALID;
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        MGLC_CUDA(cudaGetLastError());
        if (h->comm && S->direct) {
            int e = 0;
            MGLC_CUDA(cudaMemcpy(&e, S->d_err, sizeof e, cudaMemcpyDeviceToHost));
            if (e) { set_error("jacobi direct halo path: barrier %s (ranks out of step?)", (e & 1) ? "timed out" : "saw the other ping-pong array"); return MGLC_E_STATE; }
        }
    }
    return MGLC_OK;
}
// halo transport of mglc_jacobi_step on several subdomains: 1 = direct stores into the neighbours' ghost layers (default when
// the mappings came up), 0 = exchange_message, then sweep (LAP:94-103 as written)
extern "C" int mglc_jacobi_set_halo(mglc_jacobi *h, int mode) {
    if (!h || mode < 0 || mode > 1) { set_error("mglc_jacobi_set_halo: mode %d (0 exchange, 1 direct halo stores)", mode); return MGLC_E_INVALID; }
    MGLC_TRY(jac_drain(h));
    h->halo_mode = mode;
    return MGLC_OK;
}
extern "C" int mglc_jacobi_direct_halo(mglc_jacobi *h, int *available) {
    if (!h || !available) return MGLC_E_INVALID;
    int a = h->nranks > 1 ? 1 : 0;
    for (JacSub *S : h->subs) a = a && S->direct;
    *available = a;
    return MGLC_OK;
}
