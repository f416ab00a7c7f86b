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
//   k_jacobi3d_tma   the same update as a persistent TMA pipeline (the default in 3-D): one CTA per SM marches tiles of
//                128 x 16 cells along z; a producer thread streams the (130 x 18) halo'd planes into a ring of shared-memory
//                stages with cp.async.bulk.tensor (completion on mbarriers), 16 consumer warps keep z-1 / z / z+1 of their
//                4 cells in registers and take the x / y neighbours from the stage; the loads in flight no longer depend
//                on registers or occupancy
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
    CUtensorMap tmap[2]; // 3-D tiled maps of A[0], A[1] (box TSX x TSY x 1 doubles) for k_jacobi3d_tma
    CUtensorMap *tmap_dev; // the same two maps in device memory (MGLC_JACOBI_TMAP=global)
    int tma_ok;
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

template <bool HAS_F>
__global__ void __launch_bounds__(128) k_jacobi2d(Geom g, const double *__restrict__ A, const double *__restrict__ f,
                                                  double *__restrict__ B) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, 1);
    double s = A[c - 1] + A[c + 1];
    s += A[c - g.sy];
    s += A[c + g.sy];
    s += HAS_F ? f[c] : 0.0;
    B[c] = 0.25 * s;
}

// Register blocking, no barriers: a thread owns JRY consecutive rows of one x column and marches `kch` planes along
// z with the z-1 / z / z+1 values of its JRY cells in registers.  y neighbours are the thread's own registers
// (plus one row above and one below the strip per plane), x neighbours come from the adjacent lanes by shuffle
// (warp-edge lanes read them), z neighbours from the marching registers: per cell one DRAM read, one write and
// about 2/JRY extra L2 reads.
constexpr int JTX = 128;
template <bool HAS_F, int JRY>
__global__ void __launch_bounds__(JTX, JRY <= 4 ? 8 : 4) k_jacobi3d(Geom g, const double *__restrict__ A, const double *__restrict__ f,
                                                                   double *__restrict__ B, int k_lo, int k_hi, int kch) {
    const int i = 1 + blockIdx.x * JTX + threadIdx.x;
    const int j0 = 1 + blockIdx.y * JRY;
    const int k0 = k_lo + blockIdx.z * kch;
    const int k1 = min(k0 + kch - 1, k_hi);
    const int lane = threadIdx.x & 31;
    const bool active = i <= g.nx;
    const bool edge_m = lane == 0, edge_p = lane == 31 || i >= g.nx;
    const long long sy = g.sy, sz = g.sz;
    long long row[JRY];                           // rows past ny fold onto the ghost row ny+1: loaded, never stored
#pragma unroll
    for (int r = 0; r < JRY; ++r) row[r] = (long long)(min(j0 + r, g.ny + 1) - j0) * sy;
    const long long row_p = (long long)(min(j0 + JRY, g.ny + 1) - j0) * sy;
    long long c = g.idx(0, min(i, g.nx), j0, k0);
    double below[JRY], cen[JRY], up[JRY];
#pragma unroll
    for (int r = 0; r < JRY; ++r) { below[r] = A[c + row[r] - sz]; cen[r] = A[c + row[r]]; }
    for (int k = k0; k <= k1; ++k) {
#pragma unroll
        for (int r = 0; r < JRY; ++r) up[r] = A[c + row[r] + sz];
        const double ym = A[c - sy], yp = A[c + row_p];
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
            if (active && j0 + r <= g.ny) B[c + row[r]] = (1.0 / 6.0) * s;
        }
#pragma unroll
        for (int r = 0; r < JRY; ++r) { below[r] = cen[r]; cen[r] = up[r]; }
        c += sz;
    }
}

// ---- the TMA pipeline --------------------------------------------------------------------------------------------------
// Tile = TBX x TBY cells of one z plane; a stage holds the tile with its one-cell x / y rim: (TBX+2) x (TBY+2) doubles, dense.
// Work = tiles x planes, cut into z slabs of `slab` planes; inside a slab the tile-planes (tile-major, z-minor) are dealt to
// the CTAs as equal contiguous ranges, so every CTA streams the same number of planes (no last-wave tail) and all CTAs stay
// within one slab of each other in z: the rims a tile shares with its neighbours are read from DRAM once and hit the L2 after.
constexpr int TBX = 128, TBY = 16, TRY = 4;                 // tile, rows per consumer thread
constexpr int TMA_CONSUMERS = TBX * (TBY / TRY);            // 512 threads = 16 warps
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;             // + the producer warp
// XH = x rim of a stage in cells: 2 makes the first column of every box (x index OX - 2 + 128 t) start on a 16-byte boundary
constexpr int TMA_XH = 2;
constexpr int TSX = TBX + 2 * TMA_XH, TSY = TBY + 2;
constexpr int TMA_STAGE_BYTES = ((TSX * TSY * 8 + 127) / 128) * 128;
constexpr int tma_smem(int stages) { return stages * TMA_STAGE_BYTES + 2 * stages * 8 + 128; }

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

// the ranges of (tile, k) a CTA owns, slab after slab; producer and consumers walk the same sequence
struct TmaWalk {
    int ntiles, nz, slab, k_lo;
    long long pos, end;        // tile-planes of the current slab still to do: [pos, end)
    int s, hs;                 // slab index, its height
    __device__ void init(int ntiles_, int k_lo_, int k_hi_, int slab_) {
        ntiles = ntiles_; k_lo = k_lo_; nz = k_hi_ - k_lo_ + 1; slab = slab_; s = -1; pos = end = 0;
    }
    // next segment: tile t, planes ka..kb (inclusive); false when the CTA is done
    __device__ bool next(int &t, int &ka, int &kb) {
        while (pos >= end) {
            if (++s >= (nz + slab - 1) / slab) return false;
            hs = min(slab, nz - s * slab);
            const long long w = (long long)ntiles * hs;
            pos = w * blockIdx.x / gridDim.x;
            end = w * (blockIdx.x + 1) / gridDim.x;
        }
        t = (int)(pos / hs);
        const int z0 = (int)(pos - (long long)t * hs);
        const int z1 = (int)min((long long)hs, z0 + (end - pos));
        ka = k_lo + s * slab + z0;
        kb = k_lo + s * slab + z1 - 1;
        pos += z1 - z0;
        return true;
    }
};

// TMA_STAGES = 10 with one CTA per SM (169 KB of planes in flight), or 5 with two CTAs per SM
template <bool HAS_F, int TMA_STAGES>
__global__ void __launch_bounds__(TMA_THREADS, TMA_STAGES > 5 ? 1 : 2) k_jacobi3d_tma(const __grid_constant__ CUtensorMap mapP,
                                                                 const CUtensorMap *__restrict__ mapG, Geom g,
                                                                 const double *__restrict__ f, double *__restrict__ B,
                                                                 int k_lo, int k_hi, int slab) {
    const CUtensorMap *map = mapG ? mapG : &mapP;              // descriptor in global memory, or the kernel parameter
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t *full = (uint64_t *)(base + TMA_STAGES * TMA_STAGE_BYTES), *empty = full + TMA_STAGES;
    const int tiles_x = (g.nx + TBX - 1) / TBX, tiles_y = (g.ny + TBY - 1) / TBY;
    if (threadIdx.x == 0) {
        for (int q = 0; q < TMA_STAGES; ++q) { mbar_init(full + q, 1); mbar_init(empty + q, TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    TmaWalk wk;
    wk.init(tiles_x * tiles_y, k_lo, k_hi, slab);
    int t, ka, kb;
    unsigned n = 0;                                            // running plane counter: stage = n % STAGES
    if (threadIdx.x >= TMA_CONSUMERS) {
        // ---- producer: one thread streams planes ka-1 .. kb+1 of every segment into the ring
        if (threadIdx.x == TMA_CONSUMERS) {
            while (wk.next(t, ka, kb)) {
                const int x0 = (t % tiles_x) * TBX + OX - TMA_XH, y0 = (t / tiles_x) * TBY;      // cell (i0 - XH, j0 - 1)
                for (int k = ka - 1; k <= kb + 1; ++k, ++n) {
                    const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
                    mbar_wait(empty + st, ph ^ 1u);
                    mbar_expect_tx(full + st, TSX * TSY * 8);
                    tma_load_plane(base + st * TMA_STAGE_BYTES, map, full + st, x0, y0, k);
                }
            }
        }
        return;
    }
    // ---- consumers: thread = one x column, TRY consecutive rows
    const int lx = threadIdx.x % TBX, rg = threadIdx.x / TBX;
    const int so = (rg * TRY + 1) * TSX + lx + TMA_XH;         // my first cell inside a stage
    const bool lane0 = (threadIdx.x & 31) == 0;
    while (wk.next(t, ka, kb)) {
        const int i = 1 + (t % tiles_x) * TBX + lx, j0 = 1 + (t / tiles_x) * TBY + rg * TRY;
        const bool active = i <= g.nx;
        long long c = g.idx(0, min(i, g.nx), min(j0, g.ny + 1), ka);
        double below[TRY], cen[TRY], up[TRY];
        {   // plane ka-1: only my own cells, then the stage is free again
            const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
            mbar_wait(full + st, ph);
            const double *S = (const double *)(base + st * TMA_STAGE_BYTES);
#pragma unroll
            for (int r = 0; r < TRY; ++r) below[r] = S[so + r * TSX];
            __syncwarp();
            if (lane0) mbar_arrive(empty + st);
            ++n;
        }
        {   // plane ka
            const unsigned st = n % TMA_STAGES, ph = (n / TMA_STAGES) & 1u;
            mbar_wait(full + st, ph);
            const double *S = (const double *)(base + st * TMA_STAGE_BYTES);
#pragma unroll
            for (int r = 0; r < TRY; ++r) cen[r] = S[so + r * TSX];
        }
        for (int k = ka; k <= kb; ++k) {
            const unsigned st = n % TMA_STAGES;                                           // holds plane k (already waited for)
            const unsigned su = (n + 1) % TMA_STAGES, pu = ((n + 1) / TMA_STAGES) & 1u;   // plane k+1
            const double *S = (const double *)(base + st * TMA_STAGE_BYTES);
            const double *U = (const double *)(base + su * TMA_STAGE_BYTES);
            mbar_wait(full + su, pu);
#pragma unroll
            for (int r = 0; r < TRY; ++r) up[r] = U[so + r * TSX];
            const double ym = S[so - TSX], yp = S[so + TRY * TSX];
#pragma unroll
            for (int r = 0; r < TRY; ++r) {
                double s = S[so + r * TSX - 1] + S[so + r * TSX + 1];
                s += (r == 0) ? ym : cen[r - 1];
                s += (r == TRY - 1) ? yp : cen[r + 1];
                s += below[r];
                s += up[r];
                const long long cr = c + (long long)r * g.sy;
                if (active && j0 + r <= g.ny) {
                    s += HAS_F ? f[cr] : 0.0;
                    B[cr] = (1.0 / 6.0) * s;
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
static void jac_make_tmaps(JacSub *S) {
    S->tma_ok = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return;
    const cuuint64_t dims[3] = {(cuuint64_t)S->g.px, (cuuint64_t)S->g.py, (cuuint64_t)S->g.pz};
    const cuuint64_t strides[2] = {(cuuint64_t)S->g.sy * 8, (cuuint64_t)S->g.sz * 8};
    const cuuint32_t box[3] = {TSX, TSY, 1}, estr[3] = {1, 1, 1};
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
    if (cudaFuncSetAttribute(k_jacobi3d_tma<false, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem(10)) != cudaSuccess ||
        cudaFuncSetAttribute(k_jacobi3d_tma<true, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem(10)) != cudaSuccess ||
        cudaFuncSetAttribute(k_jacobi3d_tma<false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem(5)) != cudaSuccess ||
        cudaFuncSetAttribute(k_jacobi3d_tma<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem(5)) != cudaSuccess) { (void)cudaGetLastError(); return; }
    if (cudaMalloc((void **)&S->tmap_dev, 2 * sizeof(CUtensorMap)) != cudaSuccess ||
        cudaMemcpy(S->tmap_dev, S->tmap, 2 * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) { (void)cudaGetLastError(); return; }
    S->tma_ok = 1;
}
static int jacobi_tma_ctas() {            // CTAs per SM of the TMA pipeline: 1 (10 stages) or 2 (5 stages each)
    const char *e = getenv("MGLC_JACOBI_TMA_CTAS");
    return e && atoi(e) == 2 ? 2 : 1;
}
// MGLC_JACOBI_KERNEL=reg keeps the register-blocked LDG kernel (k_jacobi3d); default: the TMA pipeline
// (the three switches are read at every sweep, so a test or a tuning sweep can flip them inside one process)
static bool jacobi_use_tma() {
    const char *e = getenv("MGLC_JACOBI_KERNEL");
    return !(e && !strcmp(e, "reg"));
}
static int jacobi_slab(int nz) {
    int v = 64;
    if (const char *e = getenv("MGLC_JACOBI_SLAB")) v = std::max(1, atoi(e));
    return std::min(v, nz);
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
    for (Msg &M : S->msgs) { cudaFree(M.sbuf); cudaFree(M.rbuf); }
    cudaEvent_t evs[] = {S->ev_packed, S->ev_copied, S->ev_t0, S->ev_t1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (S->s) cudaStreamDestroy(S->s);
    delete S;
}

extern "C" int mglc_jacobi_destroy(mglc_jacobi *h) {
    if (!h) return MGLC_OK;
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
    MGLC_TRY(jac_copy(h, S, A, S->A[S->cur], false));
    MGLC_TRY(jac_copy(h, S, A_new, S->A[S->cur ^ 1], false));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}

// init(), LAP:144-166: everything 0, the top ghost layer of the top ranks 1 (rims included), A_p = A
extern "C" int mglc_jacobi_init(mglc_jacobi *h) {
    if (!h) return MGLC_E_INVALID;
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

// jacobi(A, A_new), LAP:170-182, then the roles swap (LAP:97-103 ping-pongs the two arrays)
static int jac_sweep(mglc_jacobi *h) {
    for (JacSub *S : h->subs) {
        MGLC_TRY(jac_use(S));
        const double *A = S->A[S->cur];
        double *B = S->A[S->cur ^ 1];
        if (h->ndim == 2) {
            const dim3 grid((S->n[0] + 127) / 128, S->n[1]);
            if (S->f) k_jacobi2d<true><<<grid, 128, 0, S->s>>>(S->g, A, S->f, B);
            else k_jacobi2d<false><<<grid, 128, 0, S->s>>>(S->g, A, nullptr, B);
        } else if (S->tma_ok && jacobi_use_tma()) {
            const int tiles = ((S->n[0] + TBX - 1) / TBX) * ((S->n[1] + TBY - 1) / TBY);
            const int slab = jacobi_slab(S->n[2]);
            const int per_sm = jacobi_tma_ctas();
            const int grid = (int)std::min<long long>((long long)sm_count(S->device) * per_sm, (long long)tiles * slab);
            const char *where = getenv("MGLC_JACOBI_TMAP");
            const CUtensorMap *gmap = (where && !strcmp(where, "global")) ? S->tmap_dev + S->cur : nullptr;
#define MGLC_JAC_TMA(HASF, NST, FPTR) k_jacobi3d_tma<HASF, NST><<<grid, TMA_THREADS, tma_smem(NST), S->s>>>(S->tmap[S->cur], gmap, S->g, FPTR, B, 1, S->n[2], slab)
            if (per_sm == 2) { if (S->f) MGLC_JAC_TMA(true, 5, S->f); else MGLC_JAC_TMA(false, 5, nullptr); }
            else { if (S->f) MGLC_JAC_TMA(true, 10, S->f); else MGLC_JAC_TMA(false, 10, nullptr); }
#undef MGLC_JAC_TMA
        } else {
            const int kch = jacobi_kch(S->n[2]);
            const int JRY = jacobi_jry();
            const dim3 grid((S->n[0] + JTX - 1) / JTX, (S->n[1] + JRY - 1) / JRY, (S->n[2] + kch - 1) / kch);
            const dim3 block(JTX);
            if (JRY == 8) {
                if (S->f) k_jacobi3d<true, 8><<<grid, block, 0, S->s>>>(S->g, A, S->f, B, 1, S->n[2], kch);
                else k_jacobi3d<false, 8><<<grid, block, 0, S->s>>>(S->g, A, nullptr, B, 1, S->n[2], kch);
            } else {
                if (S->f) k_jacobi3d<true, 4><<<grid, block, 0, S->s>>>(S->g, A, S->f, B, 1, S->n[2], kch);
                else k_jacobi3d<false, 4><<<grid, block, 0, S->s>>>(S->g, A, nullptr, B, 1, S->n[2], kch);
            }
        }
        S->launches += 1;
        S->cur ^= 1;
    }
    return MGLC_OK;
}

extern "C" int mglc_jacobi_exchange(mglc_jacobi *h) { if (!h) return MGLC_E_INVALID; return jac_exchange(h); }
extern "C" int mglc_jacobi_sweep(mglc_jacobi *h) { if (!h) return MGLC_E_INVALID; return jac_sweep(h); }

static int jac_step_impl(mglc_jacobi *h, int nits) {
    if (nits < 0) { set_error("mglc_jacobi_step: nits=%d", nits); return MGLC_E_INVALID; }
    for (int it = 0; it < nits; ++it) {
        MGLC_TRY(jac_exchange(h));
        MGLC_TRY(jac_sweep(h));
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
    for (JacSub *S : h->subs) { MGLC_TRY(jac_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
