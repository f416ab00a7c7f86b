// exact_kernels.cu -- kernels whose results are bit-defined (copies, order-preserving sums, IEEE
// divisions): initial(), streaming(), bounceback(), macro(), check() reductions, halo pack/unpack and the AoS<->SoA transposes of upload/download.
// Built with -fmad=false so nothing is contracted; see lbm_kernels.inl for the collision kernels.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "d3q19_mrt.inl"
#include "d3q19_thermal.inl"

namespace mglc {

__constant__ int c_ex[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
__constant__ int c_ey[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
__constant__ int c_ez[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
__constant__ int c_opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
__constant__ int c_face_pops[6][5] = {{1, 7, 9, 11, 13}, {2, 8, 10, 12, 14}, {3, 7, 8, 15, 17},
                                      {4, 9, 10, 16, 18}, {5, 11, 12, 15, 16}, {6, 13, 14, 17, 18}};

static inline dim3 grid3(int nx, int ny, int nz, int tx) { return dim3((nx + tx - 1) / tx, ny, nz); }

// ---- initial(), L3/initial.f90:46-73 -----------------------------------------------------------
__global__ void k_initial(Geom g, LbmParams p, double *__restrict__ F, double *__restrict__ rho,
                          double *__restrict__ u, double *__restrict__ v, double *__restrict__ w) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const double r0 = p.rho0;
    const double uu = (g.lid && k == g.nz) ? p.U0 : 0.0, vv = 0.0, ww = 0.0;   // top boundary, :55-61
    const long long m = g.cell(i, j, k);
    rho[m] = r0; u[m] = uu; v[m] = vv; w[m] = ww;
    const double us2 = uu * uu + vv * vv + ww * ww;
    const long long c = g.idx(0, i, j, k);
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        const double omega = (a == 0) ? 1.0 / 3.0 : (a < 7 ? 1.0 / 18.0 : 1.0 / 36.0);
        const double un = uu * (double)c_ex[a] + vv * (double)c_ey[a] + ww * (double)c_ez[a];
        F[a * g.sq + c] = r0 * omega * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
    }
}

int launch_initial(const Geom &g, const LbmParams &p, double *F, double *rho, double *u, double *v, double *w,
                   cudaStream_t s) {
    k_initial<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, p, F, rho, u, v, w);
    return 1;
}

// ---- streaming(), L3/streaming.f90:8-20: f(a,x) = f_post(a, x - e_a), halo read as it is ---------
__global__ void __launch_bounds__(128) k_streaming(Geom g, const double *__restrict__ Fpost, double *__restrict__ F) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, k);
#pragma unroll
    for (int a = 0; a < Q; ++a)
        F[a * g.sq + c] = Fpost[a * g.sq + c - c_ez[a] * g.sz - c_ey[a] * g.sy - c_ex[a]];
}

int launch_streaming(const Geom &g, const double *Fpost, double *F, cudaStream_t s) {
    k_streaming<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, Fpost, F);
    return 1;
}

// moving-lid correction of L3/bounce_back.f90:77-78 for the bounced population `a` (13 or 14)
__device__ __forceinline__ double lid_term(double val, double rho, double U0, int a) {
    const double uw = (a == 14) ? U0 : -U0;
    return val - rho / 6.0 * uw;
}

// ---- bounceback(), L3/bounce_back.f90:6-83, in place on f after streaming ------------------------
// One pass per wall face; a population entering through two walls gets the same value from both
// passes (half-way bounce-back), and the lid term is keyed on the cell (k == nz on the lid rank) so the
// reference's "z processed last" outcome holds whichever pass writes last.
__global__ void k_bounceback(Geom g, LbmParams p, const double *__restrict__ Fpost, const double *__restrict__ rho,
                             double *__restrict__ F) {
    const int face = blockIdx.y;
    if (!g.wall[face]) return;
    const int axis = face >> 1;
    const int n1 = (axis == 0) ? g.ny : g.nx, n2 = (axis == 2) ? g.ny : g.nz;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * n2) return;
    const int t1 = 1 + t % n1, t2 = 1 + t / n1;
    const int nfix = (axis == 0) ? g.nx : (axis == 1 ? g.ny : g.nz);
    const int fix = (face & 1) ? 1 : nfix;                 // +face -> last layer, -face -> first layer
    const int i = (axis == 0) ? fix : t1;
    const int j = (axis == 1) ? fix : (axis == 0 ? t1 : t2);
    const int k = (axis == 2) ? fix : t2;
    const long long c = g.idx(0, i, j, k);
    const bool on_lid = g.lid && k == g.nz;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int a = c_face_pops[face ^ 1][q];            // populations that come out of this wall
        double val = Fpost[c_opp[a] * g.sq + c];
        if (on_lid && (a == 13 || a == 14)) val = lid_term(val, rho[g.cell(i, j, k)], p.U0, a);
        F[a * g.sq + c] = val;
    }
}

int launch_bounceback(const Geom &g, const LbmParams &p, const double *Fpost, const double *rho, double *F,
                      cudaStream_t s) {
    const int big = max(max(g.nx * g.ny, g.nx * g.nz), g.ny * g.nz);
    k_bounceback<<<dim3((big + 127) / 128, 6), 128, 0, s>>>(g, p, Fpost, rho, F);
    return 1;
}

// ---- macro(), L3/macro.f90:6-25 ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_macro(Geom g, const double *__restrict__ F, double *__restrict__ rho,
                                               double *__restrict__ u, double *__restrict__ v, double *__restrict__ w) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, k);
    double f[19];
#pragma unroll
    for (int a = 0; a < Q; ++a) f[a] = F[a * g.sq + c];
    double r, uu, vv, ww;
    d3q19_macro(f, r, uu, vv, ww);
    const long long m = g.cell(i, j, k);
    rho[m] = r; u[m] = uu; v[m] = vv; w[m] = ww;
}

int launch_macro(const Geom &g, const double *F, double *rho, double *u, double *v, double *w, cudaStream_t s) {
    k_macro<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, F, rho, u, v, w);
    return 1;
}

// ---- check(), L3/check.f90:12-23: two sums (error1 has no w term, :15) + up,vp,wp <- u,v,w ----------
constexpr int CHECK_BLOCKS = 592;   // 4 x 148 SMs; fixed so the summation order is reproducible
__global__ void __launch_bounds__(256) k_check_partial(long long n, const double *__restrict__ u,
                                                       const double *__restrict__ v, const double *__restrict__ w,
                                                       double *__restrict__ up, double *__restrict__ vp,
                                                       double *__restrict__ wp, double *__restrict__ part) {
    double e1 = 0.0, e2 = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const double a = u[q], b = v[q], c = w[q];
        const double da = a - up[q], db = b - vp[q];
        e1 += da * da + db * db;
        e2 += a * a + b * b + c * c;
        up[q] = a; vp[q] = b; wp[q] = c;
    }
    __shared__ double s1[256], s2[256];
    s1[threadIdx.x] = e1; s2[threadIdx.x] = e2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 + 2 * blockIdx.x] = s1[0]; part[3 + 2 * blockIdx.x] = s2[0]; }
}
__global__ void k_check_final(int nblocks, double *__restrict__ part) {
    double e1 = 0.0, e2 = 0.0;
    for (int b = 0; b < nblocks; ++b) { e1 += part[2 + 2 * b]; e2 += part[3 + 2 * b]; }
    part[0] = e1; part[1] = e2;
}
int check_scratch_doubles() { return 4 + 4 * CHECK_BLOCKS; }
int launch_check(const Geom &g, const double *u, const double *v, const double *w, double *up, double *vp,
                 double *wp, double *part, cudaStream_t s) {
    const long long n = (long long)g.nx * g.ny * g.nz;
    k_check_partial<<<CHECK_BLOCKS, 256, 0, s>>>(n, u, v, w, up, vp, wp, part);
    k_check_final<<<1, 1, 0, s>>>(CHECK_BLOCKS, part);
    return 2;
}

// ---- halo pack / unpack, replacing the MPI derived datatypes surface_x/y/z, line_x/y/z -------------
// (L3/main.f90:52-72) and the 42 Sendrecv of L3/ex_sendrecv.f90.  Message `dir`: 0..5 = faces
// +x,-x,+y,-y,+z,-z carrying 5 populations in ascending order; 7..18 = the edge crossed by that
// population.  Buffer layout [slot][t2][t1], t1 the faster tangential index.
// cell coordinates of element (t1,t2) of message dir; recv=0: source layer in the interior,
// recv=1: destination layer in the halo of the receiver
__device__ __forceinline__ void msg_cell(const Geom &g, int dir, int recv, int t1, int t2, int &i, int &j, int &k) {
    if (dir < 6) {
        const int axis = dir >> 1, plus = !(dir & 1);
        const int nfix = (axis == 0) ? g.nx : (axis == 1 ? g.ny : g.nz);
        const int fix = recv ? (plus ? 0 : nfix + 1) : (plus ? nfix : 1);
        i = (axis == 0) ? fix : 1 + t1;
        j = (axis == 1) ? fix : (axis == 0 ? 1 + t1 : 1 + t2);
        k = (axis == 2) ? fix : 1 + t2;
    } else {
        const int e[3] = {c_ex[dir], c_ey[dir], c_ez[dir]};
        const int n[3] = {g.nx, g.ny, g.nz};
        int x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (e[d] == 0) x[d] = 1 + t1;
            else if (e[d] > 0) x[d] = recv ? 0 : n[d];
            else x[d] = recv ? n[d] + 1 : 1;
        }
        i = x[0]; j = x[1]; k = x[2];
    }
}

__global__ void k_pack(Geom g, const double *__restrict__ Fpost, int dir, int n1, int n2, int npop,
                       double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per = (long long)n1 * n2;
    if (t >= per * npop) return;
    const int slot = (int)(t / per);
    const int r = (int)(t % per);
    int i, j, k;
    msg_cell(g, dir, 0, r % n1, r / n1, i, j, k);
    const int a = (dir < 6) ? c_face_pops[dir][slot] : dir;
    buf[t] = Fpost[g.idx(a, i, j, k)];
}
__global__ void k_unpack(Geom g, double *__restrict__ Fpost, int dir, int n1, int n2, int npop,
                         const double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per = (long long)n1 * n2;
    if (t >= per * npop) return;
    const int slot = (int)(t / per);
    const int r = (int)(t % per);
    int i, j, k;
    msg_cell(g, dir, 1, r % n1, r / n1, i, j, k);
    const int a = (dir < 6) ? c_face_pops[dir][slot] : dir;
    Fpost[g.idx(a, i, j, k)] = buf[t];
}

void msg_dims(const Geom &g, int dir, int &n1, int &n2, int &npop) {
    if (dir < 6) {
        const int axis = dir >> 1;
        n1 = (axis == 0) ? g.ny : g.nx; n2 = (axis == 2) ? g.ny : g.nz; npop = 5;
    } else {
        n1 = (h_ex[dir] == 0) ? g.nx : (h_ey[dir] == 0 ? g.ny : g.nz); n2 = 1; npop = 1;
    }
}
int launch_pack(const Geom &g, const double *Fpost, int dir, double *buf, cudaStream_t s) {
    int n1, n2, npop;
    msg_dims(g, dir, n1, n2, npop);
    const long long n = (long long)n1 * n2 * npop;
    k_pack<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, Fpost, dir, n1, n2, npop, buf);
    return 1;
}
int launch_unpack(const Geom &g, double *Fpost, int dir, const double *buf, cudaStream_t s) {
    int n1, n2, npop;
    msg_dims(g, dir, n1, n2, npop);
    const long long n = (long long)n1 * n2 * npop;
    k_unpack<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, Fpost, dir, n1, n2, npop, buf);
    return 1;
}

// all messages of one exchange in ONE launch (blockIdx.y = message): 18 f messages + 6 g messages used to be 24 launches of
// a few microseconds each on either side of the NCCL call
template <bool UNPACK>
__global__ void k_pack_all(Geom g, MsgBatch mb, double *Fpost, double *Gpost) {
    const int m = blockIdx.y;
    const int dir = mb.dir[m];
    const bool isg = dir >= 20;
    const int d = isg ? dir - 20 : dir;
    int n1, n2, npop;
    if (d < 6) { const int axis = d >> 1; n1 = (axis == 0) ? g.ny : g.nx; n2 = (axis == 2) ? g.ny : g.nz; npop = isg ? 1 : 5; }
    else { n1 = (c_ex[d] == 0) ? g.nx : (c_ey[d] == 0 ? g.ny : g.nz); n2 = 1; npop = 1; }
    const long long per = (long long)n1 * n2;
    double *buf = mb.buf[m];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per * npop; t += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(t / per), r = (int)(t % per);
        int i, j, k;
        msg_cell(g, d, UNPACK ? 1 : 0, r % n1, r / n1, i, j, k);
        const int a = isg ? d + 1 : (d < 6 ? c_face_pops[d][slot] : d);
        double *lat = isg ? Gpost : Fpost;
        if (UNPACK) lat[g.idx(a, i, j, k)] = buf[t];
        else buf[t] = lat[g.idx(a, i, j, k)];
    }
}
int launch_pack_all(const Geom &g, const MsgBatch &mb, double *Fpost, double *Gpost, bool unpack, cudaStream_t s) {
    if (mb.n == 0) return 0;
    const long long face = (long long)std::max(g.nx, g.ny) * std::max(g.ny, g.nz) * 5;
    const dim3 grid((unsigned)std::min<long long>(1024, (face + 255) / 256), mb.n);
    if (unpack) k_pack_all<true><<<grid, 256, 0, s>>>(g, mb, Fpost, Gpost);
    else k_pack_all<false><<<grid, 256, 0, s>>>(g, mb, Fpost, Gpost);
    return 1;
}

// ---- halo push (transport 3): after a fused launch WITHOUT in-kernel peer stores, one launch copies the outgoing
// populations of every face / edge message (the sets of ex_sendrecv.f90:12-123, + the g population per face for thermal
// lattices) from this subdomain's boundary cells straight into the neighbours' halo cells over NVLink.  No send buffer, no
// NCCL, no unpack; the neighbour barrier (k_halo_signal / k_halo_wait) orders it like the in-kernel stores.
__global__ void k_push_halos(Geom g, const PeerTable *__restrict__ pt, const double *__restrict__ F, const double *__restrict__ G) {
    const int m = blockIdx.y;                       // 0..18: f message in direction m (6 = none); 19..24: g message of face m - 19
    const bool isg = m >= 19;
    const int d = isg ? m - 19 : m;
    if (d == 6 || !(pt->mask >> d & 1u) || (isg && !G)) return;
    if (pt->err && *pt->err) return;                // the barrier failed: nothing goes into a neighbour any more
    int n1, n2, npop;
    if (d < 6) { const int axis = d >> 1; n1 = (axis == 0) ? g.ny : g.nx; n2 = (axis == 2) ? g.ny : g.nz; npop = isg ? 1 : 5; }
    else { n1 = (c_ex[d] == 0) ? g.nx : (c_ey[d] == 0 ? g.ny : g.nz); n2 = 1; npop = 1; }
    const int e[3] = {d < 6 ? (d >> 1 == 0 ? 1 - 2 * (d & 1) : 0) : c_ex[d], d < 6 ? (d >> 1 == 1 ? 1 - 2 * (d & 1) : 0) : c_ey[d],
                      d < 6 ? (d >> 1 == 2 ? 1 - 2 * (d & 1) : 0) : c_ez[d]};
    const long long per = (long long)n1 * n2, psy = pt->sy[d], psz = pt->sz[d], psq = pt->sq[d];
    double *P = isg ? pt->G[d] : pt->F[d];
    const double *src = isg ? G : F;
    double *xs = d < 2 ? pt->XS[d] : nullptr;       // x faces: contiguous staging in the neighbour's memory (PeerTable::XS)
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < per * npop; t += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(t / per), r = (int)(t % per);
        int i, j, k;
        msg_cell(g, d, 0, r % n1, r / n1, i, j, k);
        const int a = isg ? d + 1 : (d < 6 ? c_face_pops[d][slot] : d);
        if (xs) { xs[(isg ? 5 * per : 0) + t] = src[g.idx(a, i, j, k)]; continue; }
        // the neighbour's halo cell: layer 0 / n'+1 along the axes the message crosses, my own index along the others
        const int ii = e[0] > 0 ? 0 : (e[0] < 0 ? pt->n[d][0] + 1 : i);
        const int jj = e[1] > 0 ? 0 : (e[1] < 0 ? pt->n[d][1] + 1 : j);
        const int kk = e[2] > 0 ? 0 : (e[2] < 0 ? pt->n[d][2] + 1 : k);
        P[a * psq + kk * psz + jj * psy + (ii + OX - 1)] = src[g.idx(a, i, j, k)];
    }
}
int launch_push_halos(const Geom &g, const PeerTable *pt_dev, const double *F, const double *G, cudaStream_t s) {
    const long long face = (long long)std::max(g.nx, g.ny) * std::max(g.ny, g.nz) * 5;
    const dim3 grid((unsigned)std::min<long long>(512, (face + 255) / 256), G ? 25 : 19);
    k_push_halos<<<grid, 256, 0, s>>>(g, pt_dev, F, G);
    return 1;
}

// ---- AoS (reference: population index fastest) <-> SoA transposes ----------------------------------
// A block moves TILE consecutive cells (linear reference order) through shared memory so that both the
// AoS side (19*TILE contiguous doubles) and the SoA side (TILE contiguous cells per population) coalesce.
constexpr int TILE = 128;
// NQ = populations per cell in the AoS array (19 for f / f_post, 7 for g / g_post)
__device__ __forceinline__ long long soa_cell_index(const Geom &g, long long cell, int with_halo) {
    const int ex_ = with_halo ? 2 : 0;
    const long long lx = g.nx + ex_, ly = g.ny + ex_;
    const int i = (int)(cell % lx), j = (int)((cell / lx) % ly), k = (int)(cell / (lx * ly));
    const int o = with_halo ? 0 : 1;
    return g.idx(0, i + o, j + o, k + o);
}
template <int NQ>
__global__ void __launch_bounds__(256) k_aos_to_soa(Geom g, const double *__restrict__ aos, double *__restrict__ F,
                                                    long long c0, long long ncells, int with_halo) {
    constexpr int Q = NQ;
    __shared__ double sm[TILE * Q];
    const long long t0 = (long long)blockIdx.x * TILE;
    const int nt = (int)min((long long)TILE, ncells - t0);
    for (int q = threadIdx.x; q < nt * Q; q += blockDim.x) sm[q] = aos[t0 * Q + q];
    __syncthreads();
    const int t = threadIdx.x % TILE;
    if (t < nt) {
        const long long c = soa_cell_index(g, c0 + t0 + t, with_halo);
        for (int a = threadIdx.x / TILE; a < Q; a += blockDim.x / TILE) F[a * g.sq + c] = sm[t * Q + a];
    }
}
template <int NQ>
__global__ void __launch_bounds__(256) k_soa_to_aos(Geom g, const double *__restrict__ F, double *__restrict__ aos,
                                                    long long c0, long long ncells, int with_halo) {
    constexpr int Q = NQ;
    __shared__ double sm[TILE * Q];
    const long long t0 = (long long)blockIdx.x * TILE;
    const int nt = (int)min((long long)TILE, ncells - t0);
    const int t = threadIdx.x % TILE;
    if (t < nt) {
        const long long c = soa_cell_index(g, c0 + t0 + t, with_halo);
        for (int a = threadIdx.x / TILE; a < Q; a += blockDim.x / TILE) sm[t * Q + a] = F[a * g.sq + c];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nt * Q; q += blockDim.x) aos[t0 * Q + q] = sm[q];
}
int launch_aos_to_soa(const Geom &g, int nq, const double *aos, double *F, long long c0, long long ncells, int with_halo,
                      cudaStream_t s) {
    const unsigned nb = (unsigned)((ncells + TILE - 1) / TILE);
    if (nq == 19) k_aos_to_soa<19><<<nb, 256, 0, s>>>(g, aos, F, c0, ncells, with_halo);
    else k_aos_to_soa<7><<<nb, 256, 0, s>>>(g, aos, F, c0, ncells, with_halo);
    return 1;
}
int launch_soa_to_aos(const Geom &g, int nq, const double *F, double *aos, long long c0, long long ncells, int with_halo,
                      cudaStream_t s) {
    const unsigned nb = (unsigned)((ncells + TILE - 1) / TILE);
    if (nq == 19) k_soa_to_aos<19><<<nb, 256, 0, s>>>(g, F, aos, c0, ncells, with_halo);
    else k_soa_to_aos<7><<<nb, 256, 0, s>>>(g, F, aos, c0, ncells, with_halo);
    return 1;
}

// ==================== thermal double-distribution path (B3) ====================
// initial(), B3:409-638: rho = 1, u = 0, T = 0 except the layer next to a constant-temperature y or z wall
// (B3:542-587), f = feq(rho,u), g = omegaT * T * (1 + 21/(6+paraA) e.u)
__global__ void k_th_initial(Geom g, ThermalParams tp, double *__restrict__ F, double *__restrict__ G,
                             double *__restrict__ rho, double *__restrict__ u, double *__restrict__ v,
                             double *__restrict__ w, double *__restrict__ T) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    double Tc = 0.0;
    // same order as the reference: y walls (hot side first), then z walls
    if (g.wall[3] && j == 1 && tp.bcT[3]) Tc = tp.bcT[3] == MGLC_BCT_CONST_HOT ? tp.Thot : tp.Tcold;
    if (g.wall[2] && j == g.ny && tp.bcT[2]) Tc = tp.bcT[2] == MGLC_BCT_CONST_HOT ? tp.Thot : tp.Tcold;
    if (g.wall[5] && k == 1 && tp.bcT[5]) Tc = tp.bcT[5] == MGLC_BCT_CONST_HOT ? tp.Thot : tp.Tcold;
    if (g.wall[4] && k == g.nz && tp.bcT[4]) Tc = tp.bcT[4] == MGLC_BCT_CONST_HOT ? tp.Thot : tp.Tcold;
    const double r0 = 1.0, uu = 0.0, vv = 0.0, ww = 0.0;
    const long long m = g.cell(i, j, k), c = g.idx(0, i, j, k);
    rho[m] = r0; u[m] = uu; v[m] = vv; w[m] = ww; T[m] = Tc;
    const double us2 = uu * uu + vv * vv + ww * ww;
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        const double omega = (a == 0) ? 1.0 / 3.0 : (a < 7 ? 1.0 / 18.0 : 1.0 / 36.0);
        const double un = uu * (double)c_ex[a] + vv * (double)c_ey[a] + ww * (double)c_ez[a];
        F[a * g.sq + c] = r0 * omega * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
    }
#pragma unroll
    for (int a = 0; a < QT; ++a) {
        const double omegaT = (a == 0) ? (1.0 - tp.paraA) / 7.0 : (tp.paraA + 6.0) / 42.0;
        const double unT = uu * (double)c_ex[a] + vv * (double)c_ey[a] + ww * (double)c_ez[a];
        G[a * g.sq + c] = omegaT * Tc * (1.0 + 21.0 / (6.0 + tp.paraA) * unT);
    }
}
int launch_th_initial(const Geom &g, const ThermalParams &tp, double *F, double *G, double *rho, double *u, double *v,
                      double *w, double *T, cudaStream_t s) {
    k_th_initial<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, tp, F, G, rho, u, v, w, T);
    return 1;
}

// streamingT(), B3:1075-1098
__global__ void __launch_bounds__(128) k_streamingT(Geom g, const double *__restrict__ Gpost, double *__restrict__ G) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, k);
#pragma unroll
    for (int a = 0; a < QT; ++a) G[a * g.sq + c] = Gpost[a * g.sq + c - c_ez[a] * g.sz - c_ey[a] * g.sy - c_ex[a]];
}
int launch_streamingT(const Geom &g, const double *Gpost, double *G, cudaStream_t s) {
    k_streamingT<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, Gpost, G);
    return 1;
}

// bouncebackT(), B3:1100-1210, in place on g after streamingT: one population per wall face
__global__ void k_bouncebackT(Geom g, ThermalParams tp, const double *__restrict__ Gpost, double *__restrict__ G) {
    const int face = blockIdx.y;
    if (!g.wall[face]) return;
    const int axis = face >> 1;
    const int n1 = (axis == 0) ? g.ny : g.nx, n2 = (axis == 2) ? g.ny : g.nz;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * n2) return;
    const int t1 = 1 + t % n1, t2 = 1 + t / n1;
    const int nfix = (axis == 0) ? g.nx : (axis == 1 ? g.ny : g.nz);
    const int fix = (face & 1) ? 1 : nfix;
    const int i = (axis == 0) ? fix : t1;
    const int j = (axis == 1) ? fix : (axis == 0 ? t1 : t2);
    const int k = (axis == 2) ? fix : t2;
    const long long c = g.idx(0, i, j, k);
    const int a = (face & 1) ? 2 * axis + 1 : 2 * axis + 2;      // population leaving the wall
    const int o = (face & 1) ? 2 * axis + 2 : 2 * axis + 1;
    const double raw = Gpost[o * g.sq + c];
    G[a * g.sq + c] = tp.bcT[face] ? (-raw + tp.wallT[face]) : raw;
}
int launch_bouncebackT(const Geom &g, const ThermalParams &tp, const double *Gpost, double *G, cudaStream_t s) {
    const int big = max(max(g.nx * g.ny, g.nx * g.nz), g.ny * g.nz);
    k_bouncebackT<<<dim3((big + 127) / 128, 6), 128, 0, s>>>(g, tp, Gpost, G);
    return 1;
}

// macro() with the half-force term, B3:986-1010; Fc = Fx,Fy,Fz back to back
__global__ void __launch_bounds__(128) k_th_macro(Geom g, const double *__restrict__ F, const double *__restrict__ Fc,
                                                  double *__restrict__ rho, double *__restrict__ u,
                                                  double *__restrict__ v, double *__restrict__ w) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, k), m = g.cell(i, j, k), n = (long long)g.nx * g.ny * g.nz;
    double f[19];
#pragma unroll
    for (int a = 0; a < Q; ++a) f[a] = F[a * g.sq + c];
    double r, uu, vv, ww;
    d3q19_macro_forced(f, Fc[m], Fc[n + m], Fc[2 * n + m], r, uu, vv, ww);
    rho[m] = r; u[m] = uu; v[m] = vv; w[m] = ww;
}
int launch_th_macro(const Geom &g, const double *F, const double *Fc, double *rho, double *u, double *v, double *w,
                    cudaStream_t s) {
    k_th_macro<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, F, Fc, rho, u, v, w);
    return 1;
}
// macroT(), B3:1215-1232
__global__ void __launch_bounds__(128) k_macroT(Geom g, const double *__restrict__ G, double *__restrict__ T) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j, k);
    double gq[7];
#pragma unroll
    for (int a = 0; a < QT; ++a) gq[a] = G[a * g.sq + c];
    T[g.cell(i, j, k)] = d3q7_temperature(gq);
}
int launch_macroT(const Geom &g, const double *G, double *T, cudaStream_t s) {
    k_macroT<<<grid3(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, G, T);
    return 1;
}

// check(), B3:1236-1283: four sums (errorU WITH the w term here) + previous-field update
__global__ void __launch_bounds__(256) k_th_check_partial(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                          const double *__restrict__ w, const double *__restrict__ T,
                                                          double *__restrict__ up, double *__restrict__ vp,
                                                          double *__restrict__ wp, double *__restrict__ Tp,
                                                          double *__restrict__ part) {
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const double a = u[q], b = v[q], c = w[q], t = T[q];
        const double da = a - up[q], db = b - vp[q], dc = c - wp[q];
        e[0] += da * da + db * db + dc * dc;
        e[1] += a * a + b * b + c * c;
        e[2] += fabs(t - Tp[q]);
        e[3] += fabs(t);
        up[q] = a; vp[q] = b; wp[q] = c; Tp[q] = t;
    }
    __shared__ double sm[4][256];
    for (int q = 0; q < 4; ++q) sm[q][threadIdx.x] = e[q];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) for (int q = 0; q < 4; ++q) sm[q][threadIdx.x] += sm[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 4) part[4 + 4 * blockIdx.x + threadIdx.x] = sm[threadIdx.x][0];
}
__global__ void k_th_check_final(int nblocks, double *__restrict__ part) {
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = 0; b < nblocks; ++b) for (int q = 0; q < 4; ++q) e[q] += part[4 + 4 * b + q];
    for (int q = 0; q < 4; ++q) part[q] = e[q];
}
int launch_th_check(const Geom &g, const double *u, const double *v, const double *w, const double *T, double *up, double *vp,
                    double *wp, double *Tp, double *part, cudaStream_t s) {
    const long long n = (long long)g.nx * g.ny * g.nz;
    k_th_check_partial<<<CHECK_BLOCKS, 256, 0, s>>>(n, u, v, w, T, up, vp, wp, Tp, part);
    k_th_check_final<<<1, 1, 0, s>>>(CHECK_BLOCKS, part);
    return 2;
}

// g halo messages, B3:1421-1468: face `face` carries population face+1 of the last interior layer
__global__ void k_pack_g(Geom g, const double *__restrict__ Gpost, int face, int n1, int n2, double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n1 * n2) return;
    int i, j, k;
    msg_cell(g, face, 0, (int)(t % n1), (int)(t / n1), i, j, k);
    buf[t] = Gpost[g.idx(face + 1, i, j, k)];
}
__global__ void k_unpack_g(Geom g, double *__restrict__ Gpost, int face, int n1, int n2, const double *__restrict__ buf) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n1 * n2) return;
    int i, j, k;
    msg_cell(g, face, 1, (int)(t % n1), (int)(t / n1), i, j, k);
    Gpost[g.idx(face + 1, i, j, k)] = buf[t];
}
int launch_pack_g(const Geom &g, const double *Gpost, int face, double *buf, cudaStream_t s) {
    int n1, n2, npop;
    msg_dims(g, face, n1, n2, npop);
    k_pack_g<<<(unsigned)(((long long)n1 * n2 + 255) / 256), 256, 0, s>>>(g, Gpost, face, n1, n2, buf);
    return 1;
}
int launch_unpack_g(const Geom &g, double *Gpost, int face, const double *buf, cudaStream_t s) {
    int n1, n2, npop;
    msg_dims(g, face, n1, n2, npop);
    k_unpack_g<<<(unsigned)(((long long)n1 * n2 + 255) / 256), 256, 0, s>>>(g, Gpost, face, n1, n2, buf);
    return 1;
}

__global__ void k_fill(double *p, long long n, double v) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) p[q] = v;
}
// ---- calNuRe(): volume sums  sum(w*T)  and  sum(u*u+v*v+w*w)  (B3/mpi_blocked/RaNu.F90:13-22,36-45) ----------------
// part[0..1] = the two sums of this subdomain; fixed grid and tree order, so the result is reproducible run to run
__global__ void __launch_bounds__(256) k_nure_partial(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                      const double *__restrict__ w, const double *__restrict__ T,
                                                      double *__restrict__ part) {
    double a1 = 0.0, a2 = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const double a = u[q], b = v[q], c = w[q];
        a1 += c * T[q];
        a2 += a * a + b * b + c * c;
    }
    __shared__ double s1[256], s2[256];
    s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[4 + 2 * blockIdx.x] = s1[0]; part[5 + 2 * blockIdx.x] = s2[0]; }
}
__global__ void k_nure_final(int nblocks, double *__restrict__ part) {
    double a1 = 0.0, a2 = 0.0;
    for (int b = 0; b < nblocks; ++b) { a1 += part[4 + 2 * b]; a2 += part[5 + 2 * b]; }
    part[0] = a1; part[1] = a2;
}
int launch_nure(const Geom &g, const double *u, const double *v, const double *w, const double *T, double *part, cudaStream_t s) {
    const long long n = (long long)g.nx * g.ny * g.nz;
    k_nure_partial<<<CHECK_BLOCKS, 256, 0, s>>>(n, u, v, w, T, part);
    k_nure_final<<<1, 1, 0, s>>>(CHECK_BLOCKS, part);
    return 2;
}

// ---- neighbour barrier of the direct-halo path --------------------------------------------------------------------
// After a fused launch that stored into the neighbours' halos, every subdomain raises its epoch in a flag word that
// lives in each neighbour's memory (signal); before anything reads its own halos it waits until all neighbours have
// raised theirs (wait).  Kernel completion orders the halo stores before the flag store on the same stream.  Like
// MPI_Waitall (lid3_mpi_nonblock.f90:1229) the wait lasts as long as the slowest neighbour's host takes to issue its
// launch; it is bounded (MGLC_HALO_TIMEOUT_S, default 600 s) only so that ranks that really fell out of step cannot hang
// the device for ever.  Running out of time does NOT go unnoticed: *err becomes sticky, no later launch stores into a
// neighbour's halos (PeerTable::err) and the host reports MGLC_E_STATE at every point where it synchronises.
__global__ void k_halo_signal(SyncTable t, unsigned long long epoch /* the word, see k_halo_wait */) {
    const int d = threadIdx.x;
    if (d < 19 && (t.mask >> d & 1u)) {
        __threadfence_system();
        *(volatile unsigned long long *)t.signal[d] = epoch;
    }
}
// word = 2 * epoch + ping-pong index of the lattice the launch wrote: neighbours out of step on the index are reported
__global__ void k_halo_wait(SyncTable t, unsigned long long word, int *err, unsigned long long timeout_ns) {
    const int d = threadIdx.x;
    if (d < 19 && (t.mask >> d & 1u)) {
        unsigned long long t0, seen;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (((seen = *(volatile unsigned long long *)t.wait[d]) >> 1) < (word >> 1)) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns || *(volatile int *)err) { atomicOr(err, 1); break; }
            __nanosleep(200);
        }
        if ((seen >> 1) == (word >> 1) && ((seen ^ word) & 1ull)) atomicOr(err, 2);
        __threadfence_system();
    }
}
double halo_timeout_seconds() {
    static double v = -1.0;
    if (v < 0.0) {
        v = 600.0;
        if (const char *e = getenv("MGLC_HALO_TIMEOUT_S")) { const double q = atof(e); if (q > 0.0) v = q; }
    }
    return v;
}
int launch_halo_signal(const SyncTable &t, unsigned long long epoch, cudaStream_t s) {
    k_halo_signal<<<1, 32, 0, s>>>(t, epoch);
    return 1;
}
int launch_halo_wait(const SyncTable &t, unsigned long long word, int *err, cudaStream_t s) {
    k_halo_wait<<<1, 32, 0, s>>>(t, word, err, (unsigned long long)(halo_timeout_seconds() * 1e9));
    return 1;
}

int launch_fill(double *p, long long n, double value, cudaStream_t s) {
    k_fill<<<1184, 256, 0, s>>>(p, n, value);
    return 1;
}

}  // namespace mglc
