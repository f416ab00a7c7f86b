// particles.cu -- particle-laden D2Q9 path of the reference (MPI/Micro_particles/fortran/case4/mpi_particle/,
// "P4") behind the mglc_p2d_* entry points of mglc.h: MRT fluid with a solid mask, quadratic-interpolated
// moving-boundary bounce-back on circular particles (calQ bisection), momentum-exchange force and torque summed by
// warp-shuffle reductions, spring repulsion, explicit particle kinematics, refill of uncovered nodes, and the
// 2-deep (f_post) / 3-deep (f) halo exchanges with corner squares.
//
// The reference keeps f and f_post as two persistent arrays and only ever touches the nodes its masks select
// (collision skips solid nodes, streaming skips populations that come out of a solid node, ...), so solid nodes
// carry stale values that refill and the bounce-back stencils may later read.  To stay comparable value for value
// the same two arrays with the same roles live on the device (no ping-pong), and every kernel applies the
// reference's conditions verbatim.  Built with -fmad=false: every operation is one IEEE rounding in the
// reference's order, so per-node results are bit-identical to the oracle; only the two reductions whose order the
// reference itself leaves open (rhoAvg and the per-particle force sums, OpenMP + MPI_Allreduce there) differ by rounding.
//
// Device layout: SoA F[a][j][x] for f and f_post with a 3-node rim (f_post uses 2 of it), rows padded so that node
// i = 1 is 128-byte aligned; obst / obstNew are ints in the same geometry; rho,u,v,up,vp are (nx,ny) like the host's.
// The whole problem (201 x 801 nodes in the shipped case) is L2-resident; one step is a handful of small launches.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "halo.cuh"

using namespace mglc;

namespace {

constexpr int Q9 = 9;
constexpr int RIM = 3;
__constant__ int c9x[Q9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};      // P4/commondata.F90:47-48
__constant__ int c9y[Q9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
__constant__ int c9r[Q9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};         // opposite, :49-50
static const int h9x[Q9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int h9y[Q9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

struct G2 {
    int nx, ny, px, i_start, j_start;
    long long sq;
    int wall[4];         // left (coords0 == 0), right, bottom (coords1 == 0), top
    __host__ __device__ long long idx(int a, int i, int j) const { return a * sq + (long long)(j + RIM - 1) * px + (i + OX - 1); }
    __host__ __device__ long long cell(int i, int j) const { return (long long)(i - 1) + (long long)nx * (j - 1); }
};

struct P2 {              // module commondata, P4/commondata.F90
    int N, total_nx, total_ny;
    double rho0, rhoSolid, Snu, Sq, gravity, thresholdWall, stiffWall, thresholdParticle, stiffParticle, radius0, Pi;
    // options taken from the reference's other particle scenario (case1/mpi_complete, "P1"); all zero = P4
    double Uwall, Uframe;        // moving top (+Uwall) / bottom (-Uwall) walls seen from a frame moving with Uframe, P1/fluid.F90:127-169
    int bb_linear, moving_walls; // linear-interpolated bounce-back on the particles, P1/particle_bounceback.F90:66-76
};

// particle state, N doubles each, one device buffer
enum { PX_ = 0, PY_, PU_, PV_, POM_, PRAD_, PINERTIA_, PXO_, PYO_, PUO_, PVO_, POMO_, PFX_, PFY_, PTQ_, PSX_, PSY_, PST_, PFIELDS_ };
// PFX/PFY/PTQ = wallTotalForceX/Y, totalTorque; PSX/PSY/PST = this rank's link sums before the Allreduce

enum { ST_STREAM = 1, ST_WALLBB = 2, ST_PBB = 4, ST_MACRO = 8, ST_FORCE = 16, ST_COUNT = 32 };
enum { ERR_CALQ = 1, ERR_Q = 2, ERR_OWNER = 4, ERR_INTERPENETRATION = 8, ERR_WALL = 16, ERR_REFILL = 32, ERR_BINS = 64 };

__device__ __forceinline__ bool inside(const G2 &g, int i, int j, double xc, double yc, double rad) {
    const double dx = (double)(i + g.i_start) - xc, dy = (double)(j + g.j_start) - yc;
    return (dx * dx + dy * dy) <= rad * rad;
}

// ---- particle bins ----------------------------------------------------------------------------------------------
// The reference finds "the particle that covers this node" and "the particles close to this particle" by looping over all
// cNumMax = 64 of them (P4/particle_update.F90:84-100, P4/particle_force.F90:100-130).  That loop is kept for small counts;
// beyond PB_MIN_PARTICLES the same questions are answered from a uniform grid of bins over the GLOBAL lattice: bin (bx, by)
// lists the particles whose centre lies in [bx*B, (bx+1)*B) x [by*B, (by+1)*B), with B >= 2*rmax + thresholdParticle (so every
// spring partner of a particle sits in the 3 x 3 bins around its own) and B >= rmax + 2 (so the particle covering a node, by its
// old or its new centre, sits in the 3 x 3 bins around the node's).  The answers do not depend on the order inside a bin:
// mask tests are an OR, and the spring sums and the refill run over the candidates in ascending particle index -- the
// reference's loop order -- so results stay bit-identical to the full loop.
constexpr int PB_CAP = 8;              // centres per bin: non-overlapping discs with 2*r + threshold <= B leave room for 5
constexpr int PB_MIN_PARTICLES = 257;  // below this the full loops run (MGLC_P2D_BINS=1 forces the bins: tests)
struct PB {
    int B, nbx, nby;                   // B == 0: no bins (small particle counts)
    double rmax;
    int *cnt, *list;
    __device__ int bin(int bx, int by) const { return by * nbx + bx; }
    __device__ int bx_of(double x) const { return min(nbx - 1, max(0, (int)floor(x / (double)B))); }
    __device__ int by_of(double y) const { return min(nby - 1, max(0, (int)floor(y / (double)B))); }
};
__global__ void k_p_bins_fill(PB b, P2 p, const double *__restrict__ ps, int *__restrict__ err) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.N) return;
    const int q = b.bin(b.bx_of(ps[PX_ * p.N + c]), b.by_of(ps[PY_ * p.N + c]));
    const int slot = atomicAdd(&b.cnt[q], 1);
    if (slot < PB_CAP) b.list[q * PB_CAP + slot] = c;
    else atomicOr(err, ERR_BINS);
}
// The particles that can cover a node of the row segment gi_lo..gi_hi of the global rows gj_lo..gj_hi (old or new centre: one
// node of slack), collected by the whole block into list[0..*nlist) (shared memory, room for 128).  All threads must call.
__device__ void block_candidates(const PB &b, int gi_lo, int gi_hi, int gj_lo, int gj_hi, int *list, int *nlist) {
    if (threadIdx.x == 0) *nlist = 0;
    __syncthreads();
    const double reach = b.rmax + 2.0;
    const int bx0 = b.bx_of((double)gi_lo - reach), bx1 = b.bx_of((double)gi_hi + reach);
    const int by0 = b.by_of((double)gj_lo - reach), by1 = b.by_of((double)gj_hi + reach);
    const int w = bx1 - bx0 + 1, nb = w * (by1 - by0 + 1);
    for (int q = threadIdx.x; q < nb * PB_CAP; q += blockDim.x) {
        const int bi = b.bin(bx0 + (q / PB_CAP) % w, by0 + (q / PB_CAP) / w), slot = q % PB_CAP;
        if (slot < min(b.cnt[bi], PB_CAP)) {
            const int pos = atomicAdd(nlist, 1);
            if (pos < 128) list[pos] = b.list[bi * PB_CAP + slot];
        }
    }
    __syncthreads();
}
// is global node (gi, gj) inside a particle?  list/nlist from block_candidates (nlist > 128: scan the bins instead)
__device__ __forceinline__ int covered(const PB &b, const G2 &g, int N, const double *__restrict__ ps, int i, int j, const int *list, int nlist) {
    int solid = 0;
    if (nlist <= 128) {
        for (int q = 0; q < nlist; ++q) {
            const int c = list[q];
            if (inside(g, i, j, ps[PX_ * N + c], ps[PY_ * N + c], ps[PRAD_ * N + c])) solid = 1;
        }
    } else {
        const int bx = b.bx_of((double)(i + g.i_start)), by = b.by_of((double)(j + g.j_start));
        for (int yy = max(0, by - 1); yy <= min(b.nby - 1, by + 1); ++yy)
            for (int xx = max(0, bx - 1); xx <= min(b.nbx - 1, bx + 1); ++xx) {
                const int bi = b.bin(xx, yy);
                for (int q = 0; q < min(b.cnt[bi], PB_CAP); ++q) {
                    const int c = b.list[bi * PB_CAP + q];
                    if (inside(g, i, j, ps[PX_ * N + c], ps[PY_ * N + c], ps[PRAD_ * N + c])) solid = 1;
                }
            }
    }
    return solid;
}

#include "p2d_calq.inl"
__device__ __forceinline__ int calQ(double xc, double yc, double rad, double i, double j, int alpha, double &x0, double &y0, double &q) {
    return calQ_link(xc, yc, rad, i, j, (double)c9x[alpha], (double)c9y[alpha], x0, y0, q);
}

// moving top / bottom walls, P1/fluid.F90:130-145 (with Uframe = 0: the `stationaryFrame` lines :151-168); fp5..fp8 = f_post of the node
__device__ __forceinline__ void p1_moving_walls(const G2 &g, const P2 &p, int j, double fp5, double fp6, double fp7, double fp8, double (&f)[9]) {
    if (g.wall[2] && j == 1) { f[5] = fp7 + (-p.Uwall - p.Uframe) / 6.0; f[6] = fp8 - (-p.Uwall - p.Uframe) / 6.0; }
    if (g.wall[3] && j == g.ny) { f[7] = fp5 - (p.Uwall - p.Uframe) / 6.0; f[8] = fp6 + (p.Uwall - p.Uframe) / 6.0; }
}
// interpolated bounce-back of one link: quadratic (P4/particle_bounceback.F90:65-75) or linear (P1/particle_bounceback.F90:67-75).
// fp0 = f_post(alpha, x), fp1 = f_post(alpha, x - e), fp2 = f_post(alpha, x - 2e), fr0 = f_post(r, x), fr1 = f_post(r, x - e); wall =
// ex(r) (Uc + temp1) + ey(r) (Vc + temp2).  fp2 / fr1 are only read by the quadratic form (passed as pointers to stay lazy).
__device__ __forceinline__ double pbb_value(int linear, double q, double omega, double rhoAvg, double fp0, double fp1, const double *fp2,
                                            double fr0, const double *fr1, double wall) {
    if (linear) {
        if (q < 0.5) return 2.0 * q * fp0 + (1.0 - 2.0 * q) * fp1 + 6.0 * omega * rhoAvg * wall;
        return 0.5 / q * fp0 + (1.0 - 0.50 / q) * fr0 + 3.0 * omega * rhoAvg / q * wall;
    }
    if (q < 0.5) return q * (1.0 + 2.0 * q) * fp0 + (1.0 - 4.0 * q * q) * fp1 - q * (1.0 - 2.0 * q) * (*fp2) + 6.0 * omega * rhoAvg * wall;
    return fp0 / q / (1.0 + 2.0 * q) + fr0 * (2.0 * q - 1.0) / q - (*fr1) * (2.0 * q - 1.0) / (2.0 * q + 1.0)
         + 6.0 * omega * rhoAvg / q / (1.0 + 2.0 * q) * wall;
}

__device__ __forceinline__ void d2q9_collide(const double (&f)[9], double rho, double u, double v, double Snu, double Sq, double (&fp)[9]) {
    double m[9], meq[9], mp[9];
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * f[5] + 2.0 * f[6] + 2.0 * f[7] + 2.0 * f[8];
    m[2] = 4.0 * f[0] - 2.0 * f[1] - 2.0 * f[2] - 2.0 * f[3] - 2.0 * f[4] + f[5] + f[6] + f[7] + f[8];
    m[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
    m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
    m[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
    m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
    m[7] = f[1] - f[2] + f[3] - f[4];
    m[8] = f[5] - f[6] + f[7] - f[8];
    meq[0] = rho;
    meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
    meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
    meq[3] = rho * u;
    meq[4] = -meq[3];
    meq[5] = rho * v;
    meq[6] = -meq[5];
    meq[7] = rho * (u * u - v * v);
    meq[8] = rho * (u * v);
    const double s[9] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, Snu, Snu};
#pragma unroll
    for (int a = 0; a < 9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);
    fp[0] = (mp[0] - mp[1] + mp[2]) / 9.0;
    fp[1] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[3] / 6.0 - mp[4] / 6.0 + mp[7] * 0.25;
    fp[2] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[5] / 6.0 - mp[6] / 6.0 - mp[7] * 0.25;
    fp[3] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[3] / 6.0 + mp[4] / 6.0 + mp[7] * 0.25;
    fp[4] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[5] / 6.0 + mp[6] / 6.0 - mp[7] * 0.25;
    fp[5] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 + mp[8] * 0.25;
    fp[6] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 - mp[8] * 0.25;
    fp[7] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 + mp[8] * 0.25;
    fp[8] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 - mp[8] * 0.25;
}

// ---- initial(), P4/initial.F90:79-199 (positions come from the caller) -----------------------------------------
__global__ void __launch_bounds__(128) k_p_initial(G2 g, P2 p, PB b, const double *__restrict__ ps, double *__restrict__ F, double *__restrict__ Fp,
                            int *__restrict__ obst, int *__restrict__ obstNew, double *__restrict__ rho, double *__restrict__ u,
                            double *__restrict__ v, double *__restrict__ up, double *__restrict__ vp) {
    __shared__ int list[128];
    __shared__ int nlist;
    const int i = -2 + (int)(blockIdx.x * blockDim.x + threadIdx.x), j = -2 + (int)blockIdx.y;
    if (b.B) block_candidates(b, -2 + (int)(blockIdx.x * blockDim.x) + g.i_start, -2 + (int)(blockIdx.x * blockDim.x + blockDim.x - 1) + g.i_start,
                              j + g.j_start, j + g.j_start, list, &nlist);
    if (i > g.nx + 3) return;
    const bool interior = i >= 1 && i <= g.nx && j >= 1 && j <= g.ny;
    const bool ring12 = !interior && i >= -1 && i <= g.nx + 2 && j >= -1 && j <= g.ny + 2;
    int solid = 0;
    if (i >= 0 && i <= g.nx + 1 && j >= 0 && j <= g.ny + 1) {
        if (b.B) solid = covered(b, g, p.N, ps, i, j, list, nlist);
        else
        for (int c = 0; c < p.N; ++c)
            if (inside(g, i, j, ps[PX_ * p.N + c], ps[PY_ * p.N + c], ps[PRAD_ * p.N + c])) solid = 1;
    }
    obst[g.idx(0, i, j)] = solid;
    obstNew[g.idx(0, i, j)] = 0;
    if (interior) {
        const long long m = g.cell(i, j);
        rho[m] = solid ? p.rhoSolid : p.rho0;
        u[m] = 0.0; v[m] = 0.0; up[m] = 0.0; vp[m] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < Q9; ++a) {
        const double omega = a == 0 ? 4.0 / 9.0 : (a < 5 ? 1.0 / 9.0 : 1.0 / 36.0);
        double fv = 0.0, fpv = 0.0;
        if (interior && !solid) {
            const double uu = 0.0, vv = 0.0, us2 = uu * uu + vv * vv;
            const double un = uu * (double)c9x[a] + vv * (double)c9y[a];
            fv = omega * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
        } else if (ring12) {
            fv = omega * p.rho0; fpv = omega * p.rho0;
        }
        F[g.idx(a, i, j)] = fv;
        Fp[g.idx(a, i, j)] = fpv;
    }
}

// Two sums over a whole grid of 128-thread blocks, reproducible: block partials (tree-reduced in shared memory) are
// stored by block index; the block that takes the last ticket adds them in a fixed order and writes out[0..1].
// Every thread of every block must call this.
__device__ void grid_sum2(double a, double b, double *__restrict__ partials, unsigned *__restrict__ ticket, double *__restrict__ out) {
    __shared__ double s1[128], s2[128];
    __shared__ bool last;
    const int t = threadIdx.x;
    s1[t] = a; s2[t] = b;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (t < o) { s1[t] += s1[t + o]; s2[t] += s2[t + o]; }
        __syncthreads();
    }
    const unsigned nb = gridDim.x * gridDim.y, bi = blockIdx.y * gridDim.x + blockIdx.x;
    if (t == 0) {
        partials[2 * bi] = s1[0]; partials[2 * bi + 1] = s2[0];
        __threadfence();
        last = atomicInc(ticket, nb - 1) == nb - 1;          // wraps to 0: ready for the next launch
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double x = 0.0, y = 0.0;
    for (unsigned q = t; q < nb; q += 128) { x += __ldcg(&partials[2 * q]); y += __ldcg(&partials[2 * q + 1]); }
    s1[t] = x; s2[t] = y;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (t < o) { s1[t] += s1[t + o]; s2[t] += s2[t + o]; }
        __syncthreads();
    }
    if (t == 0) { out[0] = s1[0]; out[1] = s2[0]; }
}

// collision() of the fused step: the same node update, plus the sums bounceback_particle() needs for rhoAvg
// (P4/particle_bounceback.F90:15-35: rho over the fluid nodes and their number) since this kernel already reads
// rho and obst of every node
__global__ void __launch_bounds__(128) k_p_collision_sum(G2 g, P2 p, const double *__restrict__ F, const double *__restrict__ rho,
                                                         const double *__restrict__ u, const double *__restrict__ v,
                                                         const int *__restrict__ obst, double *__restrict__ Fp,
                                                         double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                         double *__restrict__ out, int *__restrict__ rcount) {
    // a block walks rows blockIdx.y+1, +gridDim.y, ...: the number of block partials stays bounded on large lattices (the block
    // that adds them up at the end is a serial tail), one row per block on the reference's own sizes
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (rcount && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *rcount = 0;      // the step's refill list starts empty
    double sr = 0.0, sc = 0.0;
    for (int j = 1 + blockIdx.y; j <= g.ny; j += gridDim.y) {
        if (i <= g.nx && obst[g.idx(0, i, j)] == 0) {
            const long long c = g.idx(0, i, j), m = g.cell(i, j);
            double f[9], fp[9];
#pragma unroll
            for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
            const double r = rho[m];
            d2q9_collide(f, r, u[m], v[m], p.Snu, p.Sq, fp);
#pragma unroll
            for (int a = 0; a < 9; ++a) Fp[a * g.sq + c] = fp[a];
            sr += r; sc += 1.0;
        }
    }
    grid_sum2(sr, sc, partials, ticket, out);
}

// ---- collision(), P4/fluid.F90:1-72: fluid nodes only ------------
__global__ void __launch_bounds__(128) k_p_collision(G2 g, P2 p, const double *__restrict__ F, const double *__restrict__ rho,
                                                     const double *__restrict__ u, const double *__restrict__ v,
                                                     const int *__restrict__ obst, double *__restrict__ Fp) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > g.nx) return;
    if (obst[g.idx(0, i, j)] != 0) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9], fp[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
    d2q9_collide(f, rho[m], u[m], v[m], p.Snu, p.Sq, fp);
#pragma unroll
    for (int a = 0; a < 9; ++a) Fp[a * g.sq + c] = fp[a];
}

// sum of rho over the nodes where mask == 0, and their number (bounceback_particle :15-35, updateCenter :105-120):
// fixed grid, block partials summed in block order by the second kernel -> reproducible
constexpr int RED_BLOCKS = 148;
__global__ void __launch_bounds__(256) k_p_fluid_sum(G2 g, const double *__restrict__ rho, const int *__restrict__ mask,
                                                     double *__restrict__ part) {
    const long long n = (long long)g.nx * g.ny;
    double s = 0.0, cnt = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(q % g.nx), j = 1 + (int)(q / g.nx);
        if (mask[g.idx(0, i, j)] == 0) { s += rho[q]; cnt += 1.0; }
    }
    __shared__ double s1[256], s2[256];
    s1[threadIdx.x] = s; s2[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 + 2 * blockIdx.x] = s1[0]; part[3 + 2 * blockIdx.x] = s2[0]; }
}
__global__ void k_p_fluid_sum_final(int nblocks, double *__restrict__ part) {
    double s = 0.0, c = 0.0;
    for (int b = 0; b < nblocks; ++b) { s += part[2 + 2 * b]; c += part[3 + 2 * b]; }
    part[0] = s; part[1] = c;          // rhoAvg = part[0] / part[1] after the Allreduce
}

// ---- streaming + bounceback + bounceback_particle + macro + the link sums of calForce, one node per thread ---
// STAGES selects which of the reference's subroutines the launch performs (all of them in the fused step).
// ps_in and ps_sum are two views of ONE particle-state table (the host passes S->ps twice): the kernel reads only the rows
// PX_/PY_/PRAD_/POM_/PU_/PV_ through ps_in and writes only the rows PSX_/PSY_/PST_ through ps_sum, so no object is accessed
// through both pointers and the __restrict__ qualifiers hold (same for k_p_links below).
template <int STAGES>
__global__ void __launch_bounds__(128) k_p_update(G2 g, P2 p, const double *__restrict__ ps_in, double *__restrict__ ps_sum,
                                                  const double *__restrict__ Fp, double *__restrict__ F,
                                                  const int *__restrict__ obst, double *__restrict__ rho, double *__restrict__ u,
                                                  double *__restrict__ v, const double *__restrict__ rhoAvgPart, int *__restrict__ err,
                                                  int *__restrict__ nlinks) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    const bool in_range = i <= g.nx;
    const int ic = in_range ? i : g.nx;
    const long long c = g.idx(0, ic, j);
    const int N = p.N;
    const bool fluid = in_range && obst[c] == 0;
    double f[9];
    // ST_COUNT (fused step): number of links of this node that end in a solid node; k_p_links counts them down and the
    // thread that handles the last one redoes macro() for the node
    if ((STAGES & ST_COUNT) && in_range) {
        int nl = 0;
        if (fluid) {
#pragma unroll
            for (int a = 1; a < 9; ++a) nl += obst[c + c9y[a] * (long long)g.px + c9x[a]] == 1;
        }
        nlinks[g.cell(i, j)] = nl;
    }
    // streaming(), P4/fluid.F90:97-108: every interior node, skipped when the upstream node is solid
    if (in_range) {
        if (STAGES & ST_STREAM) {
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                const long long up_ = c - c9y[a] * (long long)g.px - c9x[a];
                if (obst[up_] == 0) { f[a] = Fp[a * g.sq + up_]; F[a * g.sq + c] = f[a]; }
                else f[a] = F[a * g.sq + c];
            }
        } else {
#pragma unroll
            for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
        }
        // bounceback(), P4/fluid.F90:115-161 (left, right, bottom, top: later walls overwrite the corners)
        if (STAGES & ST_WALLBB) {
            if (g.wall[0] && i == 1) { f[1] = Fp[3 * g.sq + c]; f[5] = Fp[7 * g.sq + c]; f[8] = Fp[6 * g.sq + c]; }
            if (g.wall[1] && i == g.nx) { f[3] = Fp[1 * g.sq + c]; f[6] = Fp[8 * g.sq + c]; f[7] = Fp[5 * g.sq + c]; }
            if (g.wall[2] && j == 1) { f[2] = Fp[4 * g.sq + c]; f[5] = Fp[7 * g.sq + c]; f[6] = Fp[8 * g.sq + c]; }
            if (g.wall[3] && j == g.ny) { f[4] = Fp[2 * g.sq + c]; f[7] = Fp[5 * g.sq + c]; f[8] = Fp[6 * g.sq + c]; }
            if (p.moving_walls && ((g.wall[2] && j == 1) || (g.wall[3] && j == g.ny)))
                p1_moving_walls(g, p, j, Fp[5 * g.sq + c], Fp[6 * g.sq + c], Fp[7 * g.sq + c], Fp[8 * g.sq + c], f);
            if ((g.wall[0] && i == 1) || (g.wall[1] && i == g.nx) || (g.wall[2] && j == 1) || (g.wall[3] && j == g.ny)) {
#pragma unroll
                for (int a = 1; a < 9; ++a) F[a * g.sq + c] = f[a];
            }
        }
    }
    // links from this fluid node into a particle
    double lfx = 0.0, lfy = 0.0, ltq = 0.0;
    int lc = -1;                     // particle the pending link sums belong to
    if ((STAGES & (ST_PBB | ST_FORCE)) && fluid) {
        const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
        for (int a = 1; a < 9; ++a) {            // alpha = 0 never points into a solid from a fluid node
            const int ip = i + c9x[a], jp = j + c9y[a];
            if (obst[g.idx(0, ip, jp)] != 1) continue;
            int found = 0;
            for (int cn = 0; cn < N; ++cn) {
                const double xc = ps_in[PX_ * N + cn], yc = ps_in[PY_ * N + cn], rad = ps_in[PRAD_ * N + cn];
                if (!inside(g, ip, jp, xc, yc, rad)) continue;
                found = 1;
                double x0, y0, q;
                const int rc = calQ(xc, yc, rad, (double)(i + g.i_start), (double)(j + g.j_start), a, x0, y0, q);
                if (rc) { atomicOr(err, rc); continue; }
                const double om = ps_in[POM_ * N + cn], Uc = ps_in[PU_ * N + cn], Vc = ps_in[PV_ * N + cn];
                const double temp1 = -(y0 - yc) * om, temp2 = (x0 - xc) * om;
                const int ra = c9r[a];
                const double exr = (double)c9x[ra], eyr = (double)c9y[ra];
                if (STAGES & ST_PBB) {           // P4/particle_bounceback.F90:65-75
                    const double omega = a < 5 ? 1.0 / 9.0 : 1.0 / 36.0;
                    const double fp0 = Fp[a * g.sq + c];
                    const long long c1 = c - c9y[a] * (long long)g.px - c9x[a];
                    const long long c2 = c1 - c9y[a] * (long long)g.px - c9x[a];
                    const double val = pbb_value(p.bb_linear, q, omega, rhoAvg, fp0, Fp[a * g.sq + c1], Fp + (a * g.sq + c2), Fp[ra * g.sq + c],
                                                 Fp + (ra * g.sq + c1), exr * (Uc + temp1) + eyr * (Vc + temp2));
                    f[ra] = val;
                    F[ra * g.sq + c] = val;
                }
                if (STAGES & ST_FORCE) {         // P4/particle_force.F90:60-62 (after macro in the reference; f is final here too,
                                                 // unless a later link of this node rewrites f[ra] -- handled by the second pass below)
                    if (!(STAGES & ST_PBB)) {
                        const double fpa = Fp[a * g.sq + c], fr = f[ra];
                        const double tfx = ((double)c9x[a] - Uc - temp1) * fpa - (exr - Uc - temp1) * fr;
                        const double tfy = ((double)c9y[a] - Vc - temp2) * fpa - (eyr - Vc - temp2) * fr;
                        const double ttq = (x0 - xc) * tfy - (y0 - yc) * tfx;
                        if (lc != cn && lc >= 0) {     // flush the sums of the previous particle (rare: a node between two particles)
                            atomicAdd(&ps_sum[PSX_ * N + lc], lfx); atomicAdd(&ps_sum[PSY_ * N + lc], lfy); atomicAdd(&ps_sum[PST_ * N + lc], ltq);
                            lfx = lfy = ltq = 0.0;
                        }
                        lc = cn; lfx += tfx; lfy += tfy; ltq += ttq;
                    }
                }
            }
            if (!found) atomicOr(err, ERR_OWNER);
        }
    }
    // fused launch: the link sums need the node's FINAL f (every bounce-back of this node applied), so they run as
    // a second pass over the links once the loop above is complete
    if ((STAGES & ST_FORCE) && (STAGES & ST_PBB) && fluid) {
        for (int a = 1; a < 9; ++a) {
            const int ip = i + c9x[a], jp = j + c9y[a];
            if (obst[g.idx(0, ip, jp)] != 1) continue;
            for (int cn = 0; cn < N; ++cn) {
                const double xc = ps_in[PX_ * N + cn], yc = ps_in[PY_ * N + cn], rad = ps_in[PRAD_ * N + cn];
                if (!inside(g, ip, jp, xc, yc, rad)) continue;
                double x0, y0, q;
                if (calQ(xc, yc, rad, (double)(i + g.i_start), (double)(j + g.j_start), a, x0, y0, q)) continue;
                const double om = ps_in[POM_ * N + cn], Uc = ps_in[PU_ * N + cn], Vc = ps_in[PV_ * N + cn];
                const double temp1 = -(y0 - yc) * om, temp2 = (x0 - xc) * om;
                const int ra = c9r[a];
                const double fpa = Fp[a * g.sq + c], fr = f[ra];
                const double tfx = ((double)c9x[a] - Uc - temp1) * fpa - ((double)c9x[ra] - Uc - temp1) * fr;
                const double tfy = ((double)c9y[a] - Vc - temp2) * fpa - ((double)c9y[ra] - Vc - temp2) * fr;
                const double ttq = (x0 - xc) * tfy - (y0 - yc) * tfx;
                if (lc != cn && lc >= 0) {
                    atomicAdd(&ps_sum[PSX_ * N + lc], lfx); atomicAdd(&ps_sum[PSY_ * N + lc], lfy); atomicAdd(&ps_sum[PST_ * N + lc], ltq);
                    lfx = lfy = ltq = 0.0;
                }
                lc = cn; lfx += tfx; lfy += tfy; ltq += ttq;
            }
        }
    }
    // macro(), P4/fluid.F90:164-184: fluid nodes only
    if ((STAGES & ST_MACRO) && fluid) {
        const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        const long long m = g.cell(i, j);
        rho[m] = r;
        u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / r;
        v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / r;
    }
    // momentum-exchange sums: lanes of a warp that hit the same particle combine by shuffles, one atomic per
    // (warp, particle)
    if (STAGES & ST_FORCE) {
        unsigned pending = __ballot_sync(0xffffffffu, lc >= 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int cn = __shfl_sync(0xffffffffu, lc, leader);
            const bool mine = lc == cn;
            double sx = mine ? lfx : 0.0, sy = mine ? lfy : 0.0, st = mine ? ltq : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_down_sync(0xffffffffu, sx, o);
                sy += __shfl_down_sync(0xffffffffu, sy, o);
                st += __shfl_down_sync(0xffffffffu, st, o);
            }
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&ps_sum[PSX_ * N + cn], sx); atomicAdd(&ps_sum[PSY_ * N + cn], sy); atomicAdd(&ps_sum[PST_ * N + cn], st);
            }
            if (mine) lc = -1;
            pending = __ballot_sync(0xffffffffu, lc >= 0);
        }
    }
}


// ---- fused step: the node pass (streaming + bounceback + macro + link count) laid out for bandwidth ------------------
// Same result as k_p_update<STREAM | WALLBB | MACRO | COUNT>, node for node and bit for bit.  The nine masks around the node
// are loaded once (the upstream node of population a is the downstream node of its opposite, so they also give the link
// count), then the nine populations are loaded in one batch through selected addresses: f_post of the upstream node, or --
// where that node is solid and streaming() leaves f untouched (P4/fluid.F90:97-108) -- the node's own old f.  No load waits
// for a mask test in between, and every population is stored (an untouched one is stored back unchanged).
__global__ void __launch_bounds__(128) k_p_node(G2 g, P2 p, const double *Fp, double *F, const int *__restrict__ obst, double *__restrict__ rho,
                                                double *__restrict__ u, double *__restrict__ v, int *__restrict__ nlinks) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j);
    int ob[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) ob[a] = obst[c - c9y[a] * (long long)g.px - c9x[a]];
    const bool fluid = ob[0] == 0;
    int nl = 0;
    if (fluid) {
#pragma unroll
        for (int a = 1; a < 9; ++a) nl += ob[c9r[a]] == 1;          // obst(x + e_a) = the upstream mask of the opposite direction
    }
    nlinks[g.cell(i, j)] = nl;
    double f[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        const double *src = ob[a] == 0 ? Fp + (a * g.sq + c - c9y[a] * (long long)g.px - c9x[a]) : F + (a * g.sq + c);
        f[a] = *src;
    }
    // bounceback(), P4/fluid.F90:115-161 (left, right, bottom, top: later walls overwrite the corners)
    if (g.wall[0] && i == 1) { f[1] = Fp[3 * g.sq + c]; f[5] = Fp[7 * g.sq + c]; f[8] = Fp[6 * g.sq + c]; }
    if (g.wall[1] && i == g.nx) { f[3] = Fp[1 * g.sq + c]; f[6] = Fp[8 * g.sq + c]; f[7] = Fp[5 * g.sq + c]; }
    if (g.wall[2] && j == 1) { f[2] = Fp[4 * g.sq + c]; f[5] = Fp[7 * g.sq + c]; f[6] = Fp[8 * g.sq + c]; }
    if (g.wall[3] && j == g.ny) { f[4] = Fp[2 * g.sq + c]; f[7] = Fp[5 * g.sq + c]; f[8] = Fp[6 * g.sq + c]; }
    if (p.moving_walls && ((g.wall[2] && j == 1) || (g.wall[3] && j == g.ny)))
        p1_moving_walls(g, p, j, Fp[5 * g.sq + c], Fp[6 * g.sq + c], Fp[7 * g.sq + c], Fp[8 * g.sq + c], f);
#pragma unroll
    for (int a = 0; a < 9; ++a) F[a * g.sq + c] = f[a];
    if (fluid) {                                                     // macro(), P4/fluid.F90:164-184
        const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        const long long m = g.cell(i, j);
        rho[m] = r;
        u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / r;
        v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / r;
    }
}

// ---- fused step: bounceback_particle() + the link sums of calForce(), one block per particle ------------------------
// The node kernel (k_p_update<STREAM|WALLBB|MACRO|COUNT>) leaves in nlinks the number of links of every fluid node that
// end in a solid node.  Here block cn scans the bounding box of particle cn for those links (fluid node -> node inside
// THIS particle; no search over the other particles), gives every link its own thread (prefix sum over the per-node
// counts -> queue in shared memory), and per link does what P4/particle_bounceback.F90:52-75 and
// P4/particle_force.F90:52-66 do: calQ bisection, quadratic-interpolated bounce-back into f(r(alpha)), momentum exchange
// with that final value.  The thread that finishes a node's last link redoes macro() for it (P4/fluid.F90:164-184).
// Link sums are reduced in a fixed order (thread-strided partial sums, then a shared-memory tree): reproducible, no atomics.
constexpr int LK_T = 256;
__global__ void __launch_bounds__(LK_T) k_p_links(G2 g, P2 p, const double *__restrict__ ps, double *__restrict__ ps_sum,
                                                  const double *__restrict__ Fp, double *F, const int *__restrict__ obst,
                                                  int *nlinks, double *rho, double *u, double *v,
                                                  const double *__restrict__ rhoAvgPart, int *__restrict__ err) {
    __shared__ int queue[LK_T * 8];
    __shared__ int wsum[LK_T / 32];
    __shared__ double r1[LK_T], r2[LK_T], r3[LK_T];
    const int cn = blockIdx.x, N = p.N, t = threadIdx.x;
    const double xc = ps[PX_ * N + cn], yc = ps[PY_ * N + cn], rad = ps[PRAD_ * N + cn];
    const double om = ps[POM_ * N + cn], Uc = ps[PU_ * N + cn], Vc = ps[PV_ * N + cn];
    const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
    // local index range of the fluid nodes that can touch the particle (one node beyond its extent), clipped to the interior
    const int li0 = max(1, (int)floor(xc - rad) - 1 - g.i_start), li1 = min(g.nx, (int)ceil(xc + rad) + 1 - g.i_start);
    const int lj0 = max(1, (int)floor(yc - rad) - 1 - g.j_start), lj1 = min(g.ny, (int)ceil(yc + rad) + 1 - g.j_start);
    const int bw = li1 - li0 + 1, bh = lj1 - lj0 + 1;
    const int nnodes = (bw > 0 && bh > 0) ? bw * bh : 0;
    if (nnodes == 0) {                   // the particle does not reach this subdomain (most of them, on a decomposed lattice)
        if (t == 0) { ps_sum[PSX_ * N + cn] = 0.0; ps_sum[PSY_ * N + cn] = 0.0; ps_sum[PST_ * N + cn] = 0.0; }
        return;
    }
    double lfx = 0.0, lfy = 0.0, ltq = 0.0;
    // one link: bisection, bounce-back, momentum exchange, and macro() when it was the node's last
    auto do_link = [&](int e) {
        const int a = e & 15, nn = e >> 4;
        const int i = li0 + nn % bw, j = lj0 + nn / bw;
        const long long c = g.idx(0, i, j), m = g.cell(i, j);
        double x0, y0, q;
        const int rc = calQ(xc, yc, rad, (double)(i + g.i_start), (double)(j + g.j_start), a, x0, y0, q);
        if (rc) atomicOr(err, rc);
        else {
            const double temp1 = -(y0 - yc) * om, temp2 = (x0 - xc) * om;
            const int ra = c9r[a];
            const double exr = (double)c9x[ra], eyr = (double)c9y[ra];
            const double omega = a < 5 ? 1.0 / 9.0 : 1.0 / 36.0;
            const double fp0 = Fp[a * g.sq + c];
            const long long c1 = c - c9y[a] * (long long)g.px - c9x[a];
            const long long c2 = c1 - c9y[a] * (long long)g.px - c9x[a];
            const double val = pbb_value(p.bb_linear, q, omega, rhoAvg, fp0, Fp[a * g.sq + c1], Fp + (a * g.sq + c2), Fp[ra * g.sq + c],
                                         Fp + (ra * g.sq + c1), exr * (Uc + temp1) + eyr * (Vc + temp2));
            F[ra * g.sq + c] = val;
            // momentum exchange over the link, P4/particle_force.F90:60-62, with the node's final f(r(alpha)) = val
            const double tfx = ((double)c9x[a] - Uc - temp1) * fp0 - (exr - Uc - temp1) * val;
            const double tfy = ((double)c9y[a] - Vc - temp2) * fp0 - (eyr - Vc - temp2) * val;
            lfx += tfx; lfy += tfy; ltq += (x0 - xc) * tfy - (y0 - yc) * tfx;
        }
        __threadfence();
        if (atomicSub(&nlinks[m], 1) == 1) {     // every link of this node is in: macro(), P4/fluid.F90:164-184
            __threadfence();
            double f[9];
#pragma unroll
            for (int b = 0; b < 9; ++b) f[b] = __ldcg(&F[b * g.sq + c]);
            const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
            rho[m] = r;
            u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / r;
            v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / r;
        }
    };
    // pass 1: scan the box LK_T nodes at a time and queue every link (prefix sum of the per-node counts gives each
    // link a fixed slot); pass 2: one link per thread.  The queue is drained early only if a box holds more links
    // than it has room for (a particle far larger than the shipped radius).
    int queued = 0;                                            // uniform across the block
    for (int base = 0; base < nnodes; base += LK_T) {
        const int n = base + t;
        int bits = 0, cnt = 0;
        if (n < nnodes) {
            const int i = li0 + n % bw, j = lj0 + n / bw;
            const long long c = g.idx(0, i, j);
            if (obst[c] == 0) {
#pragma unroll
                for (int a = 1; a < 9; ++a) {
                    const int ip = i + c9x[a], jp = j + c9y[a];
                    if (obst[c + c9y[a] * (long long)g.px + c9x[a]] == 1 && inside(g, ip, jp, xc, yc, rad)) { bits |= 1 << a; ++cnt; }
                }
            }
        }
        int inc = cnt;                                         // inclusive prefix sum of cnt over the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if ((t & 31) >= o) inc += y; }
        __syncthreads();                                       // wsum of the previous chunk has been read
        if ((t & 31) == 31) wsum[t >> 5] = inc;
        __syncthreads();
        int off = inc - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < LK_T / 32; ++w) { if (w < (t >> 5)) off += wsum[w]; total += wsum[w]; }
        if (queued + total > LK_T * 8) {                       // no room: drain first
            for (int l = t; l < queued; l += LK_T) do_link(queue[l]);
            __syncthreads();
            queued = 0;
        }
        off += queued;
        for (int a = 1; a < 9; ++a) if (bits & (1 << a)) queue[off++] = (n << 4) | a;
        queued += total;
    }
    __syncthreads();
    for (int l = t; l < queued; l += LK_T) do_link(queue[l]);
    r1[t] = lfx; r2[t] = lfy; r3[t] = ltq;
    __syncthreads();
    for (int o = LK_T / 2; o > 0; o >>= 1) {
        if (t < o) { r1[t] += r1[t + o]; r2[t] += r2[t + o]; r3[t] += r3[t + o]; }
        __syncthreads();
    }
    if (t == 0) { ps_sum[PSX_ * N + cn] = r1[0]; ps_sum[PSY_ * N + cn] = r2[0]; ps_sum[PST_ * N + cn] = r3[0]; }
}

// ---- the per-particle tail of calForce (P4/particle_force.F90:95-190) and the kinematics of updateCenter
// (P4/particle_update.F90:25-50).  Every rank integrates every particle from the same Allreduced sums, which
// yields what the reference's owner-computes + masked Allreduce (message_particle.F90:470-497) leaves everywhere.
enum { PT_FORCES = 1, PT_ADVANCE = 2 };
__global__ void k_p_particles(P2 p, double *ps, const double *__restrict__ rhoAvgPart, int what, int *__restrict__ err) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, N = p.N;
    const bool on = c < N;
    if ((what & PT_FORCES) && on) {
        const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
        const double rad = ps[PRAD_ * N + c], xc = ps[PX_ * N + c], yc = ps[PY_ * N + c];
        double forceScale = p.Pi * (rad * rad) * (p.rhoSolid - rhoAvg) * p.gravity / p.stiffParticle;
        double Fxij = 0.0, Fyij = 0.0;
        for (int c2 = 0; c2 < N; ++c2) {
            if (c2 == c) continue;
            const double x2 = ps[PX_ * N + c2], y2 = ps[PY_ * N + c2], r2 = ps[PRAD_ * N + c2];
            const double dij = sqrt((xc - x2) * (xc - x2) + (yc - y2) * (yc - y2));
            if (dij >= (rad + r2 + p.thresholdParticle)) {
            } else if (dij >= (rad + r2)) {
                const double t = (dij - rad - r2 - p.thresholdParticle) / p.thresholdParticle;
                Fxij = Fxij + forceScale * (t * t) * (xc - x2) / dij;
                Fyij = Fyij + forceScale * (t * t) * (yc - y2) / dij;
            } else atomicOr(err, ERR_INTERPENETRATION);
        }
        double Fwx = 0.0, Fwy = 0.0;
        forceScale = p.Pi * (rad * rad) * (p.rhoSolid - p.rho0) * p.gravity / p.stiffWall;
        double dw = yc - rad - 1.0;
        if (dw < 0) atomicOr(err, ERR_WALL);
        else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwy = Fwy + forceScale * (t * t); }
        dw = xc - rad - 1.0;
        if (dw < 0) atomicOr(err, ERR_WALL);
        else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx + forceScale * (t * t); }
        dw = (double)p.total_nx - xc - rad;
        if (dw < 0) atomicOr(err, ERR_WALL);
        else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx - forceScale * (t * t); }
        ps[PFX_ * N + c] = ps[PSX_ * N + c] + Fxij + Fwx;
        ps[PFY_ * N + c] = ps[PSY_ * N + c] - (p.rhoSolid - rhoAvg) * p.Pi * (p.radius0 * p.radius0) * p.gravity + Fyij + Fwy;
        ps[PTQ_ * N + c] = ps[PST_ * N + c];
    }
    if ((what & PT_ADVANCE) && on) {
        const double xo = ps[PX_ * N + c], yo = ps[PY_ * N + c], Uo = ps[PU_ * N + c], Vo = ps[PV_ * N + c], oo = ps[POM_ * N + c];
        ps[PXO_ * N + c] = xo; ps[PYO_ * N + c] = yo; ps[PUO_ * N + c] = Uo; ps[PVO_ * N + c] = Vo; ps[POMO_ * N + c] = oo;
        const double ax = ps[PFX_ * N + c] / p.Pi / (p.radius0 * p.radius0) / p.rhoSolid;
        const double ay = ps[PFY_ * N + c] / p.Pi / (p.radius0 * p.radius0) / p.rhoSolid;
        const double aOmega = ps[PTQ_ * N + c] / ps[PINERTIA_ * N + c];       // 0.5*rhoSolid*Pi*radius**4 from the host's libm
        ps[PU_ * N + c] = Uo + ax;
        ps[PV_ * N + c] = Vo + ay;
        ps[POM_ * N + c] = oo + aOmega;
        ps[PX_ * N + c] = xo + Uo + 0.5 * ax;
        ps[PY_ * N + c] = yo + Vo + 0.5 * ay;
    }
}

// PT_FORCES for any number of particles: one thread per particle, spring partners from the 3 x 3 bins around its own (every
// particle within 2*rmax + thresholdParticle lies there).  The contributing partners (a handful) are added in ascending
// index, which is the reference's loop order, so the sums round exactly like the full loop of k_p_particles.
__global__ void __launch_bounds__(64) k_p_forces_binned(PB b, P2 p, double *ps, const double *__restrict__ rhoAvgPart, int *__restrict__ err) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, N = p.N;
    if (c >= N) return;
    const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
    const double rad = ps[PRAD_ * N + c], xc = ps[PX_ * N + c], yc = ps[PY_ * N + c];
    double forceScale = p.Pi * (rad * rad) * (p.rhoSolid - rhoAvg) * p.gravity / p.stiffParticle;
    constexpr int MAXP = 12;
    int pid[MAXP];
    double ptx[MAXP], pty[MAXP];
    int np_ = 0;
    const int bx = b.bx_of(xc), by = b.by_of(yc);
    for (int yy = max(0, by - 1); yy <= min(b.nby - 1, by + 1); ++yy)
        for (int xx = max(0, bx - 1); xx <= min(b.nbx - 1, bx + 1); ++xx) {
            const int bi = b.bin(xx, yy);
            for (int q = 0; q < min(b.cnt[bi], PB_CAP); ++q) {
                const int c2 = b.list[bi * PB_CAP + q];
                if (c2 == c) continue;
                const double x2 = ps[PX_ * N + c2], y2 = ps[PY_ * N + c2], r2 = ps[PRAD_ * N + c2];
                const double dij = sqrt((xc - x2) * (xc - x2) + (yc - y2) * (yc - y2));
                if (dij >= (rad + r2 + p.thresholdParticle)) {
                } else if (dij >= (rad + r2)) {
                    const double t = (dij - rad - r2 - p.thresholdParticle) / p.thresholdParticle;
                    if (np_ < MAXP) { pid[np_] = c2; ptx[np_] = forceScale * (t * t) * (xc - x2) / dij; pty[np_] = forceScale * (t * t) * (yc - y2) / dij; ++np_; }
                    else atomicOr(err, ERR_BINS);
                } else atomicOr(err, ERR_INTERPENETRATION);
            }
        }
    double Fxij = 0.0, Fyij = 0.0;
    for (int prev = -1, k = 0; k < np_; ++k) {          // ascending partner index
        int best = -1;
        for (int q = 0; q < np_; ++q) if (pid[q] > prev && (best < 0 || pid[q] < pid[best])) best = q;
        Fxij = Fxij + ptx[best]; Fyij = Fyij + pty[best];
        prev = pid[best];
    }
    double Fwx = 0.0, Fwy = 0.0;
    forceScale = p.Pi * (rad * rad) * (p.rhoSolid - p.rho0) * p.gravity / p.stiffWall;
    double dw = yc - rad - 1.0;
    if (dw < 0) atomicOr(err, ERR_WALL);
    else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwy = Fwy + forceScale * (t * t); }
    dw = xc - rad - 1.0;
    if (dw < 0) atomicOr(err, ERR_WALL);
    else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx + forceScale * (t * t); }
    dw = (double)p.total_nx - xc - rad;
    if (dw < 0) atomicOr(err, ERR_WALL);
    else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx - forceScale * (t * t); }
    ps[PFX_ * N + c] = ps[PSX_ * N + c] + Fxij + Fwx;
    ps[PFY_ * N + c] = ps[PSY_ * N + c] - (p.rhoSolid - rhoAvg) * p.Pi * (p.radius0 * p.radius0) * p.gravity + Fyij + Fwy;
    ps[PTQ_ * N + c] = ps[PST_ * N + c];
}

// Fused-step form of the two passes above for one block: a warp per particle evaluates the pair terms of the spring
// force 32 at a time and adds them in ascending order of the partner index (the reference's loop order, so the sums
// round identically); after a barrier the kinematics run one particle per thread.
__global__ void __launch_bounds__(1024) k_p_particles_block(P2 p, double *ps, const double *__restrict__ rhoAvgPart, int *__restrict__ err) {
    const int N = p.N, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
    for (int c = threadIdx.x >> 5; c < N; c += nw) {
        const double rad = ps[PRAD_ * N + c], xc = ps[PX_ * N + c], yc = ps[PY_ * N + c];
        double forceScale = p.Pi * (rad * rad) * (p.rhoSolid - rhoAvg) * p.gravity / p.stiffParticle;
        double Fxij = 0.0, Fyij = 0.0;
        for (int base = 0; base < N; base += 32) {
            const int c2 = base + lane;
            double tx = 0.0, ty = 0.0;
            int kind = 0;
            if (c2 < N && c2 != c) {
                const double x2 = ps[PX_ * N + c2], y2 = ps[PY_ * N + c2], r2 = ps[PRAD_ * N + c2];
                const double dij = sqrt((xc - x2) * (xc - x2) + (yc - y2) * (yc - y2));
                if (dij >= (rad + r2 + p.thresholdParticle)) {
                } else if (dij >= (rad + r2)) {
                    const double t = (dij - rad - r2 - p.thresholdParticle) / p.thresholdParticle;
                    tx = forceScale * (t * t) * (xc - x2) / dij;
                    ty = forceScale * (t * t) * (yc - y2) / dij;
                    kind = 1;
                } else kind = 2;
            }
            if (__ballot_sync(0xffffffffu, kind == 2) && lane == 0) atomicOr(err, ERR_INTERPENETRATION);
            unsigned m = __ballot_sync(0xffffffffu, kind == 1);
            while (m) {
                const int l = __ffs(m) - 1;
                Fxij = Fxij + __shfl_sync(0xffffffffu, tx, l);
                Fyij = Fyij + __shfl_sync(0xffffffffu, ty, l);
                m &= m - 1;
            }
        }
        if (lane == 0) {
            double Fwx = 0.0, Fwy = 0.0;
            forceScale = p.Pi * (rad * rad) * (p.rhoSolid - p.rho0) * p.gravity / p.stiffWall;
            double dw = yc - rad - 1.0;
            if (dw < 0) atomicOr(err, ERR_WALL);
            else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwy = Fwy + forceScale * (t * t); }
            dw = xc - rad - 1.0;
            if (dw < 0) atomicOr(err, ERR_WALL);
            else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx + forceScale * (t * t); }
            dw = (double)p.total_nx - xc - rad;
            if (dw < 0) atomicOr(err, ERR_WALL);
            else if (dw < p.thresholdWall) { const double t = (dw - p.thresholdWall) / p.thresholdWall; Fwx = Fwx - forceScale * (t * t); }
            ps[PFX_ * N + c] = ps[PSX_ * N + c] + Fxij + Fwx;
            ps[PFY_ * N + c] = ps[PSY_ * N + c] - (p.rhoSolid - rhoAvg) * p.Pi * (p.radius0 * p.radius0) * p.gravity + Fyij + Fwy;
            ps[PTQ_ * N + c] = ps[PST_ * N + c];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        const double xo = ps[PX_ * N + c], yo = ps[PY_ * N + c], Uo = ps[PU_ * N + c], Vo = ps[PV_ * N + c], oo = ps[POM_ * N + c];
        ps[PXO_ * N + c] = xo; ps[PYO_ * N + c] = yo; ps[PUO_ * N + c] = Uo; ps[PVO_ * N + c] = Vo; ps[POMO_ * N + c] = oo;
        const double ax = ps[PFX_ * N + c] / p.Pi / (p.radius0 * p.radius0) / p.rhoSolid;
        const double ay = ps[PFY_ * N + c] / p.Pi / (p.radius0 * p.radius0) / p.rhoSolid;
        const double aOmega = ps[PTQ_ * N + c] / ps[PINERTIA_ * N + c];
        ps[PU_ * N + c] = Uo + ax;
        ps[PV_ * N + c] = Vo + ay;
        ps[POM_ * N + c] = oo + aOmega;
        ps[PX_ * N + c] = xo + Uo + 0.5 * ax;
        ps[PY_ * N + c] = yo + Vo + 0.5 * ay;
    }
}

// updateCenter, P4/particle_update.F90:80-103: rebuild the mask from the new centres (rim included), reset the
// macroscopic fields of solid nodes
__global__ void __launch_bounds__(128) k_p_mask(G2 g, P2 p, PB b, const double *__restrict__ ps, int *__restrict__ obstNew,
                                                double *__restrict__ rho, double *__restrict__ u, double *__restrict__ v) {
    __shared__ int list[128];
    __shared__ int nlist;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x), j = (int)blockIdx.y;
    if (b.B) block_candidates(b, (int)(blockIdx.x * blockDim.x) + g.i_start, (int)(blockIdx.x * blockDim.x + blockDim.x - 1) + g.i_start,
                              j + g.j_start, j + g.j_start, list, &nlist);
    if (i > g.nx + 1) return;
    int solid = 0;
    if (b.B) solid = covered(b, g, p.N, ps, i, j, list, nlist);
    else
    for (int c = 0; c < p.N; ++c)
        if (inside(g, i, j, ps[PX_ * p.N + c], ps[PY_ * p.N + c], ps[PRAD_ * p.N + c])) solid = 1;
    obstNew[g.idx(0, i, j)] = solid;
    if (solid && i >= 1 && i <= g.nx && j >= 1 && j <= g.ny) {
        const long long m = g.cell(i, j);
        rho[m] = p.rhoSolid; u[m] = 0.0; v[m] = 0.0;
    }
}


// updateCenter of the fused step: the mask rebuild of k_p_mask, plus (a) the sums for the second rhoAvg
// (P4/particle_update.F90:105-120: rho over the nodes that are fluid in obstNew, after the reset above), (b) the check that
// k_p_links consumed every link the node kernel counted (a link into a solid node no particle owns: ERR_OWNER).
// A block first collects the particles whose extent reaches its row segment, then tests its nodes against those only.
__global__ void __launch_bounds__(128) k_p_mask_sum(G2 g, P2 p, PB b, const double *__restrict__ ps, const int *__restrict__ obst,
                                                    int *__restrict__ obstNew, double *__restrict__ rho, double *__restrict__ u,
                                                    double *__restrict__ v, const int *__restrict__ nlinks, int *__restrict__ err,
                                                    double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                    double *__restrict__ out, int *__restrict__ rlist, int *__restrict__ rcount, int rcap) {
    __shared__ int list[128];
    __shared__ int nlist;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x), N = p.N;
    const double gi_lo = (double)((int)(blockIdx.x * blockDim.x) + g.i_start), gi_hi = gi_lo + (double)(blockDim.x - 1);
    double sr = 0.0, sc = 0.0;
    // a block owns a band of consecutive rows (one row on the reference's own sizes); with bins the particles that can reach
    // the band are collected once for all its rows
    const int band = (g.ny + 2 + (int)gridDim.y - 1) / (int)gridDim.y;
    const int jb0 = (int)blockIdx.y * band, jb1 = min(g.ny + 1, jb0 + band - 1);
    if (b.B) block_candidates(b, (int)gi_lo, (int)gi_hi, jb0 + g.j_start, jb1 + g.j_start, list, &nlist);
    for (int j = jb0; j <= jb1; ++j) {
        if (!b.B) {
            __syncthreads();                                                   // the previous row's list has been read
            if (threadIdx.x == 0) nlist = 0;
            __syncthreads();
            const double gj = (double)(j + g.j_start);
            for (int c = threadIdx.x; c < N; c += blockDim.x) {
                const double xc = ps[PX_ * N + c], yc = ps[PY_ * N + c], rad = ps[PRAD_ * N + c];
                // conservative: one node of slack on every side of the particle's bounding box
                if (fabs(gj - yc) <= rad + 1.0 && xc + rad + 1.0 >= gi_lo && xc - rad - 1.0 <= gi_hi) {
                    const int pos = atomicAdd(&nlist, 1);
                    if (pos < 128) list[pos] = c;
                }
            }
            __syncthreads();
        }
        if (i <= g.nx + 1) {
            int solid = 0;
            if (b.B) solid = covered(b, g, N, ps, i, j, list, nlist);
            else if (nlist <= 128) {
                for (int q = 0; q < nlist; ++q) {
                    const int c = list[q];
                    if (inside(g, i, j, ps[PX_ * N + c], ps[PY_ * N + c], ps[PRAD_ * N + c])) solid = 1;
                }
            } else {
                for (int c = 0; c < N; ++c)
                    if (inside(g, i, j, ps[PX_ * N + c], ps[PY_ * N + c], ps[PRAD_ * N + c])) solid = 1;
            }
            obstNew[g.idx(0, i, j)] = solid;
            if (i >= 1 && i <= g.nx && j >= 1 && j <= g.ny) {
                const long long m = g.cell(i, j);
                if (solid) { rho[m] = p.rhoSolid; u[m] = 0.0; v[m] = 0.0; }
                else { sr += rho[m]; sc += 1.0; }
                const int was = obst[g.idx(0, i, j)];
                if (was == 0 && nlinks[m] != 0) atomicOr(err, ERR_OWNER);
                if (was == 1 && !solid) {                                      // uncovered by the move: k_p_refill's work list
                    const int pos = atomicAdd(rcount, 1);
                    if (pos < rcap) rlist[pos] = (int)m; else atomicOr(err, ERR_REFILL);
                }
            }
        }
    }
    grid_sum2(sr, sc, partials, ticket, out);
}

// updateCenter, P4/particle_update.F90:122-203: a solid node that became fluid is refilled by 3-point extrapolation
// along the lattice direction closest to the outward normal, its momentum moments reset to the wall velocity
__device__ void p_refill_node(const G2 &g, const P2 &p, const PB &b, const double *__restrict__ ps, double *__restrict__ F, double *__restrict__ rho,
                              double *__restrict__ u, double *__restrict__ v, const double *__restrict__ rhoAvgPart, int *__restrict__ err, int i, int j);
__global__ void __launch_bounds__(128) k_p_refill(G2 g, P2 p, PB b, const double *__restrict__ ps, const int *__restrict__ obst,
                                                  const int *__restrict__ obstNew, double *__restrict__ F, double *__restrict__ rho,
                                                  double *__restrict__ u, double *__restrict__ v, const double *__restrict__ rhoAvgPart,
                                                  int *__restrict__ err, const int *__restrict__ rlist, const int *__restrict__ rcount) {
    // rlist == nullptr: every interior node is tested (one node per thread).  Otherwise the launch is a fixed grid that works off
    // the list of uncovered nodes k_p_mask_sum collected (a few per particle), so no second pass over the lattice is needed.
    if (rlist) {
        const int n = *rcount;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
            const int m = rlist[q];
            p_refill_node(g, p, b, ps, F, rho, u, v, rhoAvgPart, err, 1 + m % g.nx, 1 + m / g.nx);
        }
        return;
    }
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j);
    if (!(obst[c] == 1 && obstNew[c] == 0)) return;
    p_refill_node(g, p, b, ps, F, rho, u, v, rhoAvgPart, err, i, j);
}
__device__ void p_refill_node(const G2 &g, const P2 &p, const PB &b, const double *__restrict__ ps, double *__restrict__ F, double *__restrict__ rho,
                              double *__restrict__ u, double *__restrict__ v, const double *__restrict__ rhoAvgPart, int *__restrict__ err, int i, int j) {
    const long long c = g.idx(0, i, j);
    const int N = p.N;
    const double rhoAvg = rhoAvgPart[0] / rhoAvgPart[1];
    int found = 0;
    // the refill from particle cn (whose OLD extent covered the node), P4/particle_update.F90:130-200
    auto refill_from = [&](int cn) {
        found = 1;
        const double xc = ps[PX_ * N + cn], yc = ps[PY_ * N + cn];
        const double dx = (double)(i + g.i_start) - xc, dy = (double)(j + g.j_start) - yc;
        double outNormal = 0.0;
        int ec = 0;
        for (int a = 1; a < 9; ++a) {
            const double tempNormal = (dx * (double)c9x[a] + dy * (double)c9y[a]) / sqrt(dx * dx + dy * dy);
            if (tempNormal > outNormal) { outNormal = tempNormal; ec = a; }
        }
        if (ec == 0) { atomicOr(err, ERR_REFILL); return; }
        const long long s1 = c9y[ec] * (long long)g.px + c9x[ec];
        double f[9], m[9];
#pragma unroll
        for (int a = 0; a < 9; ++a)
            f[a] = 3.0 * F[a * g.sq + c + s1] - 3.0 * F[a * g.sq + c + 2 * s1] + F[a * g.sq + c + 3 * s1];
        m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * f[5] + 2.0 * f[6] + 2.0 * f[7] + 2.0 * f[8];
        m[2] = 4.0 * f[0] - 2.0 * f[1] - 2.0 * f[2] - 2.0 * f[3] - 2.0 * f[4] + f[5] + f[6] + f[7] + f[8];
        m[3] = rhoAvg * (ps[PU_ * N + cn] - ((double)(j + g.j_start) - yc) * ps[POM_ * N + cn]);
        m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
        m[5] = rhoAvg * (ps[PV_ * N + cn] + ((double)(i + g.i_start) - xc) * ps[POM_ * N + cn]);
        m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
        m[7] = f[1] - f[2] + f[3] - f[4];
        m[8] = f[5] - f[6] + f[7] - f[8];
        f[0] = (m[0] - m[1] + m[2]) / 9.0;
        f[1] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 + m[3] / 6.0 - m[4] / 6.0 + m[7] * 0.25;
        f[2] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 + m[5] / 6.0 - m[6] / 6.0 - m[7] * 0.25;
        f[3] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 - m[3] / 6.0 + m[4] / 6.0 + m[7] * 0.25;
        f[4] = m[0] / 9.0 - m[1] / 36.0 - m[2] / 18.0 - m[5] / 6.0 + m[6] / 6.0 - m[7] * 0.25;
        f[5] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 + m[3] / 6.0 + m[4] / 12.0 + m[5] / 6.0 + m[6] / 12.0 + m[8] * 0.25;
        f[6] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 - m[3] / 6.0 - m[4] / 12.0 + m[5] / 6.0 + m[6] / 12.0 - m[8] * 0.25;
        f[7] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 - m[3] / 6.0 - m[4] / 12.0 - m[5] / 6.0 - m[6] / 12.0 + m[8] * 0.25;
        f[8] = m[0] / 9.0 + m[1] / 18.0 + m[2] / 36.0 + m[3] / 6.0 + m[4] / 12.0 - m[5] / 6.0 - m[6] / 12.0 - m[8] * 0.25;
#pragma unroll
        for (int a = 0; a < 9; ++a) F[a * g.sq + c] = f[a];
        const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        const long long mm = g.cell(i, j);
        rho[mm] = r;
        u[mm] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / r;
        v[mm] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / r;
    };
    if (!b.B) {
        for (int cn = 0; cn < N; ++cn)
            if (inside(g, i, j, ps[PXO_ * N + cn], ps[PYO_ * N + cn], ps[PRAD_ * N + cn])) refill_from(cn);
    } else {
        // candidates from the 3 x 3 bins around the node, taken in ascending particle index (the reference's loop order)
        const int bx = b.bx_of((double)(i + g.i_start)), by = b.by_of((double)(j + g.j_start));
        for (int prev = -1;;) {
            int next = N;
            for (int yy = max(0, by - 1); yy <= min(b.nby - 1, by + 1); ++yy)
                for (int xx = max(0, bx - 1); xx <= min(b.nbx - 1, bx + 1); ++xx) {
                    const int bi = b.bin(xx, yy);
                    for (int q = 0; q < min(b.cnt[bi], PB_CAP); ++q) {
                        const int cn = b.list[bi * PB_CAP + q];
                        if (cn > prev && cn < next && inside(g, i, j, ps[PXO_ * N + cn], ps[PYO_ * N + cn], ps[PRAD_ * N + cn])) next = cn;
                    }
                }
            if (next == N) break;
            refill_from(next);
            prev = next;
        }
    }
    if (!found) atomicOr(err, ERR_REFILL);
}

// check(), P4/fluid.F90:187-221: fluid nodes only
__global__ void __launch_bounds__(256) k_p_check(G2 g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ up,
                                                 double *__restrict__ vp, const int *__restrict__ obst, double *__restrict__ part) {
    const long long n = (long long)g.nx * g.ny;
    double e1 = 0.0, e2 = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(q % g.nx), j = 1 + (int)(q / g.nx);
        if (obst[g.idx(0, i, j)] != 0) continue;
        const double a = u[q], b = v[q];
        e1 += (a - up[q]) * (a - up[q]) + (b - vp[q]) * (b - vp[q]);
        e2 += a * a + b * b;
        up[q] = a; vp[q] = b;
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

// halo messages, P4/message_send_all.F90: a (ni x nj) block of nodes, all 9 populations, buffer [a][tj][ti]
__global__ void k_p_pack(G2 g, const double *__restrict__ A, int i0, int j0, int ni, int nj, double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ni * nj * Q9) return;
    const int a = t / (ni * nj), r = t % (ni * nj);
    buf[t] = A[g.idx(a, i0 + r % ni, j0 + r / ni)];
}
__global__ void k_p_unpack(G2 g, double *__restrict__ A, int i0, int j0, int ni, int nj, const double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ni * nj * Q9) return;
    const int a = t / (ni * nj), r = t % (ni * nj);
    A[g.idx(a, i0 + r % ni, j0 + r / ni)] = buf[t];
}

// host <-> device layout changes for f / f_post (AoS with rim) and obst
__global__ void k_p_aos_to_soa(G2 g, const double *__restrict__ aos, int rim, double *__restrict__ A) {
    const int i = 1 - rim + (int)(blockIdx.x * blockDim.x + threadIdx.x), j = 1 - rim + (int)blockIdx.y;
    if (i > g.nx + rim) return;
    const long long cell = (long long)(i + rim - 1) + (long long)(g.nx + 2 * rim) * (j + rim - 1);
    for (int a = 0; a < Q9; ++a) A[g.idx(a, i, j)] = aos[cell * Q9 + a];
}
__global__ void k_p_soa_to_aos(G2 g, const double *__restrict__ A, int rim, double *__restrict__ aos) {
    const int i = 1 - rim + (int)(blockIdx.x * blockDim.x + threadIdx.x), j = 1 - rim + (int)blockIdx.y;
    if (i > g.nx + rim) return;
    const long long cell = (long long)(i + rim - 1) + (long long)(g.nx + 2 * rim) * (j + rim - 1);
    for (int a = 0; a < Q9; ++a) aos[cell * Q9 + a] = A[g.idx(a, i, j)];
}
__global__ void k_p_obst_io(G2 g, int *__restrict__ host_layout, int *__restrict__ dev, int to_device) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x), j = (int)blockIdx.y;
    if (i > g.nx + 1) return;
    const long long h = (long long)i + (long long)(g.nx + 2) * j;
    if (to_device) dev[g.idx(0, i, j)] = host_layout[h];
    else host_layout[h] = dev[g.idx(0, i, j)];
}

struct Sub {
    int nx, ny, coords[2], nbr[8];   // right, left, top, bottom, tr, tl, bl, br  (-1 = MPI_PROC_NULL)
    int device;
    G2 g;
    double *F, *Fp, *rho, *u, *v, *up, *vp, *ps, *part, *stage;
    int *obst, *obstNew, *err, *istage;
    int *nlinks;                     // fused step: links into solid nodes still to be bounced back, per interior node
    double *gpart;                   // fused step: per-block partial sums of grid_sum2 (2 per block)
    unsigned *ticket;                // [0] collision+sum launch, [1] mask+sum launch
    int *rlist, *rcount, rcap;       // fused step: nodes uncovered by the particles' move (k_p_mask_sum -> k_p_refill)
    int launches_per_step, no_graph;
    int *obst0;                      // the allocation obst pointed to at create time (graph parity)
    cudaGraphExec_t gexec[2];        // one captured step per obst/obstNew parity (single-subdomain handles)
    cudaStream_t s;
    cudaEvent_t ev_packed, ev_copied, ev_t0, ev_t1;
    Msg msgs[16];                    // 0..7 f_post (depth 2), 8..15 f (depth 3)
    int org[16][6];                  // per message: send i0,j0, recv i0,j0, ni, nj
    long long launches;
    PB bins;                         // particle bins (B == 0: not in use)
};

}  // namespace

struct mglc_p2d {
    mglc_p2d_desc d;
    P2 p;
    int dims[2], nranks;
    std::vector<Sub *> subs;
    std::vector<Port> ports;
    mglc_comm *comm;
    std::vector<double> host_ps;     // staging for particle state
};

// ---- decomposition, P4/mpi_starts.F90 --------------------------------------------------------------------------------
extern "C" int mglc_p2d_dims_create(int nranks, int total_nx, int total_ny, int dims[2]) {
    if (nranks < 1 || !dims) return MGLC_E_INVALID;
    float diff = (float)(total_nx + total_ny) * (float)nranks;      // MPI_Dims_create_2d, :160-180 (default real)
    dims[0] = nranks; dims[1] = 1;
    for (int i = 1; i <= nranks; ++i)
        for (int j = 1; j <= nranks; ++j)
            if (i * j == nranks) {
                const float message = (float)(i - 1) * (float)total_ny + (float)(j - 1) * (float)total_nx;
                if (message < diff) { diff = message; dims[0] = i; dims[1] = j; }
            }
    return MGLC_OK;
}

extern "C" int mglc_p2d_desc_init(mglc_p2d_desc *d, int nparticles) {
    if (!d || nparticles < 0) return MGLC_E_INVALID;
    memset(d, 0, sizeof *d);
    const double l0 = 1.0 / 100.0, t0 = 5.0 / 10000.0;               // P4/commondata.F90:8-9
    d->total_nx = 201; d->total_ny = 801; d->nparticles = nparticles;
    d->rho0 = 1.0; d->rhoSolid = 1.01; d->viscosity = 0.05; d->radius0 = 20.0 / 2.0;
    d->thresholdWall = 6.0; d->stiffWall = 0.02; d->thresholdParticle = 6.0; d->stiffParticle = 0.08;
    d->gravity = 980.0 * (t0 * t0) / l0;
    return MGLC_OK;
}

static int cart2(const int dims[2], int c0, int c1) {
    if (c0 < 0 || c0 >= dims[0] || c1 < 0 || c1 >= dims[1]) return -1;
    return c0 * dims[1] + c1;
}

static void p_free_sub(Sub *S) {
    if (!S) return;
    cudaSetDevice(S->device);
    if (S->s) cudaStreamSynchronize(S->s);
    double *bufs[] = {S->F, S->Fp, S->rho, S->u, S->v, S->up, S->vp, S->ps, S->part, S->stage};
    for (double *b : bufs) cudaFree(b);
    cudaFree(S->obst); cudaFree(S->obstNew); cudaFree(S->err); cudaFree(S->istage);
    cudaFree(S->nlinks); cudaFree(S->gpart); cudaFree(S->ticket); cudaFree(S->rlist); cudaFree(S->rcount);
    cudaFree(S->bins.cnt); cudaFree(S->bins.list);
    for (cudaGraphExec_t e : S->gexec) if (e) cudaGraphExecDestroy(e);
    for (Msg &M : S->msgs) { cudaFree(M.sbuf); cudaFree(M.rbuf); }
    cudaEvent_t evs[] = {S->ev_packed, S->ev_copied, S->ev_t0, S->ev_t1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (S->s) cudaStreamDestroy(S->s);
    delete S;
}

extern "C" int mglc_p2d_destroy(mglc_p2d *h) {
    if (!h) return MGLC_OK;
    for (Sub *S : h->subs) p_free_sub(S);
    delete h;
    return MGLC_OK;
}

static int p_make_sub(mglc_p2d *h, int rank, int device, Sub **out) {
    Sub *S = new Sub();
    memset(S, 0, sizeof *S);
    S->device = device;
    S->coords[0] = rank / h->dims[1]; S->coords[1] = rank % h->dims[1];
    int is, js;
    mglc_decompose_1d(h->d.total_nx, S->coords[0], h->dims[0], &S->nx, &is);
    mglc_decompose_1d(h->d.total_ny, S->coords[1], h->dims[1], &S->ny, &js);
    const int c0 = S->coords[0], c1 = S->coords[1];
    const int nb[8] = {cart2(h->dims, c0 + 1, c1), cart2(h->dims, c0 - 1, c1), cart2(h->dims, c0, c1 + 1), cart2(h->dims, c0, c1 - 1),
                       cart2(h->dims, c0 + 1, c1 + 1), cart2(h->dims, c0 - 1, c1 + 1), cart2(h->dims, c0 - 1, c1 - 1), cart2(h->dims, c0 + 1, c1 - 1)};
    memcpy(S->nbr, nb, sizeof nb);
    G2 &g = S->g;
    g.nx = S->nx; g.ny = S->ny; g.i_start = is; g.j_start = js;
    g.px = ((S->nx + OX + RIM + 15) / 16) * 16;
    g.sq = (long long)g.px * (S->ny + 2 * RIM);
    g.wall[0] = c0 == 0; g.wall[1] = c0 == h->dims[0] - 1; g.wall[2] = c1 == 0; g.wall[3] = c1 == h->dims[1] - 1;
    auto fail = [&](int rc) { p_free_sub(S); return rc; };
    if (S->nx < 2 * RIM || S->ny < 2 * RIM) { set_error("mglc_p2d_create: subdomain %dx%d is thinner than the 3-node halo allows", S->nx, S->ny); return fail(MGLC_E_INVALID); }
    if (cudaSetDevice(device) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaStreamCreateWithFlags(&S->s, cudaStreamNonBlocking) != cudaSuccess) return fail(MGLC_E_CUDA);
    if (cudaEventCreateWithFlags(&S->ev_packed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&S->ev_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&S->ev_t0) != cudaSuccess || cudaEventCreate(&S->ev_t1) != cudaSuccess) return fail(MGLC_E_CUDA);
    const size_t lat = (size_t)Q9 * g.sq * sizeof(double), fld = (size_t)S->nx * S->ny * sizeof(double);
    const int N = std::max(1, h->p.N);
    if (cudaMalloc((void **)&S->F, lat) != cudaSuccess || cudaMalloc((void **)&S->Fp, lat) != cudaSuccess ||
        cudaMalloc((void **)&S->stage, (size_t)Q9 * (S->nx + 2 * RIM) * (S->ny + 2 * RIM) * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&S->obst, (size_t)g.sq * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&S->obstNew, (size_t)g.sq * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&S->istage, (size_t)(S->nx + 2) * (S->ny + 2) * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&S->err, sizeof(int)) != cudaSuccess || cudaMalloc((void **)&S->ps, (size_t)PFIELDS_ * N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&S->part, (size_t)(2 + 2 * 1024) * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); return fail(MGLC_E_NOMEM); }
    const size_t nblk = (size_t)((S->nx + 2 + 127) / 128) * (S->ny + 2);          // the larger of the two summing grids
    if (cudaMalloc((void **)&S->nlinks, (size_t)S->nx * S->ny * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&S->gpart, 2 * nblk * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&S->ticket, 2 * sizeof(unsigned)) != cudaSuccess) { (void)cudaGetLastError(); return fail(MGLC_E_NOMEM); }
    S->rcap = (int)std::min<long long>((long long)S->nx * S->ny, std::max<long long>(65536, (long long)S->nx * S->ny / 4));
    if (cudaMalloc((void **)&S->rlist, (size_t)S->rcap * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&S->rcount, sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); return fail(MGLC_E_NOMEM); }
    cudaMemsetAsync(S->rcount, 0, sizeof(int), S->s);
    cudaMemsetAsync(S->nlinks, 0, (size_t)S->nx * S->ny * sizeof(int), S->s);
    cudaMemsetAsync(S->ticket, 0, 2 * sizeof(unsigned), S->s);
    S->obst0 = S->obst;
    double **flds[] = {&S->rho, &S->u, &S->v, &S->up, &S->vp};
    for (double **f : flds) { if (cudaMalloc((void **)f, fld) != cudaSuccess) return fail(MGLC_E_NOMEM); cudaMemsetAsync(*f, 0, fld, S->s); }
    cudaMemsetAsync(S->F, 0, lat, S->s); cudaMemsetAsync(S->Fp, 0, lat, S->s);
    cudaMemsetAsync(S->obst, 0, (size_t)g.sq * sizeof(int), S->s); cudaMemsetAsync(S->obstNew, 0, (size_t)g.sq * sizeof(int), S->s);
    cudaMemsetAsync(S->err, 0, sizeof(int), S->s); cudaMemsetAsync(S->ps, 0, (size_t)PFIELDS_ * N * sizeof(double), S->s);
    cudaMemsetAsync(S->part, 0, (size_t)(2 + 2 * 1024) * sizeof(double), S->s);
    // halo plan: message m (0..7) of depth d; receiver's message m comes from the neighbour in the opposite direction
    const int opp[8] = {1, 0, 3, 2, 6, 7, 4, 5};
    for (int set = 0; set < 2; ++set) {
        const int d = set == 0 ? 2 : 3;
        const int nx = S->nx, ny = S->ny;
        // send origin / recv origin / extents per direction (P4/message_send_all.F90)
        const int sp[8][6] = {
            {nx - d + 1, 1, -d + 1, 1, d, ny},                 // right:  columns nx-d+1..nx -> -d+1..0
            {1, 1, nx + 1, 1, d, ny},                          // left:   columns 1..d       -> nx+1..nx+d
            {1, ny - d + 1, 1, -d + 1, nx, d},                 // top
            {1, 1, 1, ny + 1, nx, d},                          // bottom
            {nx - d + 1, ny - d + 1, -d + 1, -d + 1, d, d},    // top-right square  -> my bottom-left rim
            {1, ny - d + 1, nx + 1, -d + 1, d, d},             // top-left
            {1, 1, nx + 1, ny + 1, d, d},                      // bottom-left
            {nx - d + 1, 1, -d + 1, ny + 1, d, d}};            // bottom-right
        for (int m = 0; m < 8; ++m) {
            Msg &M = S->msgs[8 * set + m];
            memcpy(S->org[8 * set + m], sp[m], sizeof sp[m]);
            M.dir = 8 * set + m; M.send_to = S->nbr[m]; M.recv_from = S->nbr[opp[m]]; M.skip = 0;
            const long long cnt = (long long)sp[m][4] * sp[m][5] * Q9;
            M.send_count = M.send_to >= 0 ? cnt : 0;
            M.recv_count = M.recv_from >= 0 ? cnt : 0;
            if (M.send_count && cudaMalloc((void **)&M.sbuf, M.send_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
            if (M.recv_count && cudaMalloc((void **)&M.rbuf, M.recv_count * sizeof(double)) != cudaSuccess) return fail(MGLC_E_NOMEM);
        }
    }
    if (cudaStreamSynchronize(S->s) != cudaSuccess) return fail(MGLC_E_CUDA);
    *out = S;
    return MGLC_OK;
}

static int p_new(mglc_p2d **out, const mglc_p2d_desc *d, const int dims_or_zero[2], int nranks) {
    if (!out || !d || nranks < 1) { set_error("mglc_p2d_create: bad arguments"); return MGLC_E_INVALID; }
    if (d->total_nx < 6 || d->total_ny < 6 || d->nparticles < 0 || !(d->viscosity > 0.0)) { set_error("mglc_p2d_create: bad descriptor"); return MGLC_E_INVALID; }
    MGLC_TRY(require_gpu());
    mglc_p2d *h = new mglc_p2d();
    h->d = *d; h->nranks = nranks; h->comm = nullptr;
    if (dims_or_zero && dims_or_zero[0] > 0) { h->dims[0] = dims_or_zero[0]; h->dims[1] = dims_or_zero[1]; }
    else mglc_p2d_dims_create(nranks, d->total_nx, d->total_ny, h->dims);
    if (h->dims[0] * h->dims[1] != nranks) { set_error("mglc_p2d_create: dims do not multiply to nranks"); delete h; return MGLC_E_INVALID; }
    P2 &p = h->p;
    p.N = d->nparticles; p.total_nx = d->total_nx; p.total_ny = d->total_ny;
    p.rho0 = d->rho0; p.rhoSolid = d->rhoSolid; p.gravity = d->gravity; p.thresholdWall = d->thresholdWall; p.stiffWall = d->stiffWall;
    p.thresholdParticle = d->thresholdParticle; p.stiffParticle = d->stiffParticle; p.radius0 = d->radius0;
    p.Uwall = d->Uwall; p.Uframe = d->Uframe; p.bb_linear = d->bb_linear != 0; p.moving_walls = d->moving_walls != 0;
    p.Pi = 4.0 * atan(1.0);                                          // P4/commondata.F90:3
    const double tauf = 3.0 * d->viscosity + 0.5;                    // :33-34
    p.Snu = 1.0 / tauf;
    p.Sq = 8.0 * (2.0 * tauf - 1.0) / (8.0 * tauf - 1.0);
    h->host_ps.assign((size_t)PFIELDS_ * std::max(1, p.N), 0.0);
    *out = h;
    return MGLC_OK;
}

extern "C" int mglc_p2d_create(mglc_p2d **out, const mglc_p2d_desc *d, const int dims_or_zero[2], int nranks, int rank, int device,
                               mglc_comm *comm_or_null) {
    if (nranks > 1 && !comm_or_null) { set_error("mglc_p2d_create: %d ranks need a communicator (or use mglc_p2d_create_local)", nranks); return MGLC_E_INVALID; }
    if (rank < 0 || rank >= nranks) return MGLC_E_INVALID;
    mglc_p2d *h = nullptr;
    MGLC_TRY(p_new(&h, d, dims_or_zero, nranks));
    h->comm = comm_or_null;
    Sub *S = nullptr;
    int rc = p_make_sub(h, rank, device, &S);
    if (rc) { delete h; return rc; }
    h->subs.push_back(S);
    *out = h;
    return MGLC_OK;
}
extern "C" int mglc_p2d_create_local(mglc_p2d **out, const mglc_p2d_desc *d, const int dims_or_zero[2], int nranks, const int *devices_or_null) {
    mglc_p2d *h = nullptr;
    MGLC_TRY(p_new(&h, d, dims_or_zero, nranks));
    for (int r = 0; r < nranks; ++r) {
        Sub *S = nullptr;
        int rc = p_make_sub(h, r, devices_or_null ? devices_or_null[r] : 0, &S);
        if (rc) { mglc_p2d_destroy(h); return rc; }
        h->subs.push_back(S);
    }
    for (Sub *a : h->subs)
        for (Sub *b : h->subs)
            if (a->device != b->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->device, b->device);
                if (can) { cudaSetDevice(a->device); cudaDeviceEnablePeerAccess(b->device, 0); (void)cudaGetLastError(); }
            }
    for (Sub *S : h->subs) h->ports.push_back(Port{S->device, S->s, S->ev_packed, S->ev_copied, S->msgs, 16});
    *out = h;
    return MGLC_OK;
}

static int p_use(Sub *S) { MGLC_CUDA(cudaSetDevice(S->device)); return MGLC_OK; }
static int p_sub(mglc_p2d *h, int r, Sub **S) {
    if (!h || r < 0 || r >= (int)h->subs.size()) { set_error("mglc_p2d: bad handle or local index %d", r); return MGLC_E_INVALID; }
    *S = h->subs[r];
    return p_use(*S);
}
#define P_EACH(h, S) for (Sub * S : (h)->subs)

extern "C" int mglc_p2d_nlocal(mglc_p2d *h, int *n) { if (!h || !n) return MGLC_E_INVALID; *n = (int)h->subs.size(); return MGLC_OK; }
extern "C" int mglc_p2d_info(mglc_p2d *h, int r, int dims[2], int ln[2], int start[2], int coords[2], int nbr[8]) {
    if (!h || r < 0 || r >= (int)h->subs.size()) return MGLC_E_INVALID;
    Sub *S = h->subs[r];
    if (dims) { dims[0] = h->dims[0]; dims[1] = h->dims[1]; }
    if (ln) { ln[0] = S->nx; ln[1] = S->ny; }
    if (start) { start[0] = S->g.i_start; start[1] = S->g.j_start; }
    if (coords) { coords[0] = S->coords[0]; coords[1] = S->coords[1]; }
    if (nbr) memcpy(nbr, S->nbr, sizeof S->nbr);
    return MGLC_OK;
}

// ---- particle bins (see PB): sized from the radii whenever they are set, refilled whenever the centres move ---------------
static int p_setup_bins(mglc_p2d *h) {
    const int N = h->p.N;
    const bool want = N >= PB_MIN_PARTICLES || (N > 0 && getenv("MGLC_P2D_BINS") != nullptr);
    double rmax = 0.0;
    for (int c = 0; c < N; ++c) rmax = std::max(rmax, h->host_ps[(size_t)PRAD_ * N + c]);
    int B = 16;
    while ((double)B < std::max(2.0 * rmax + h->p.thresholdParticle, rmax + 2.0)) B *= 2;
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        PB &b = S->bins;
        const int nbx = h->p.total_nx / B + 2, nby = h->p.total_ny / B + 2;
        if (want && b.B == B && b.nbx == nbx && b.nby == nby) { b.rmax = rmax; continue; }
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        cudaFree(b.cnt); cudaFree(b.list);
        b = PB{};
        for (cudaGraphExec_t &e : S->gexec) if (e) { cudaGraphExecDestroy(e); e = nullptr; }      // captured with the old bins
        if (!want) continue;
        b.B = B; b.nbx = nbx; b.nby = nby; b.rmax = rmax;
        MGLC_CUDA(cudaMalloc((void **)&b.cnt, (size_t)nbx * nby * sizeof(int)));
        MGLC_CUDA(cudaMalloc((void **)&b.list, (size_t)nbx * nby * PB_CAP * sizeof(int)));
    }
    return MGLC_OK;
}
static int p_fill_bins(mglc_p2d *h, Sub *S) {
    const PB &b = S->bins;
    if (!b.B) return MGLC_OK;
    MGLC_CUDA(cudaMemsetAsync(b.cnt, 0, (size_t)b.nbx * b.nby * sizeof(int), S->s));
    k_p_bins_fill<<<(h->p.N + 127) / 128, 128, 0, S->s>>>(b, h->p, S->ps, S->err);
    S->launches += 1;
    return MGLC_OK;
}

// ---- particle state (replicated on every subdomain, like the reference's module arrays) ------------------------------
static int p_push_particles(mglc_p2d *h) {
    const int N = h->p.N;
    if (N == 0) return MGLC_OK;
    MGLC_TRY(p_setup_bins(h));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        MGLC_CUDA(cudaMemcpyAsync(S->ps, h->host_ps.data(), (size_t)PFIELDS_ * N * sizeof(double), cudaMemcpyHostToDevice, S->s));
        MGLC_TRY(p_fill_bins(h, S));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
    }
    return MGLC_OK;
}
static int p_pull_particles(mglc_p2d *h) {
    const int N = h->p.N;
    if (N == 0) return MGLC_OK;
    Sub *S = h->subs[0];
    MGLC_TRY(p_use(S));
    MGLC_CUDA(cudaMemcpyAsync(h->host_ps.data(), S->ps, (size_t)PFIELDS_ * N * sizeof(double), cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
extern "C" int mglc_p2d_set_particles(mglc_p2d *h, const double *x, const double *y, const double *U, const double *V,
                                      const double *omega, const double *radius) {
    if (!h) return MGLC_E_INVALID;
    const int N = h->p.N;
    MGLC_TRY(p_pull_particles(h));
    double *ps = h->host_ps.data();
    for (int c = 0; c < N; ++c) {
        if (x) ps[PX_ * N + c] = x[c];
        if (y) ps[PY_ * N + c] = y[c];
        if (U) ps[PU_ * N + c] = U[c];
        if (V) ps[PV_ * N + c] = V[c];
        if (omega) ps[POM_ * N + c] = omega[c];
        if (radius) {
            ps[PRAD_ * N + c] = radius[c];
            ps[PINERTIA_ * N + c] = 0.5 * h->p.rhoSolid * h->p.Pi * pow(radius[c], 4.0);     // P4/particle_update.F90:40
        }
    }
    return p_push_particles(h);
}
extern "C" int mglc_p2d_get_particles(mglc_p2d *h, double *x, double *y, double *U, double *V, double *omega, double *Fx, double *Fy,
                                      double *torque) {
    if (!h) return MGLC_E_INVALID;
    const int N = h->p.N;
    MGLC_TRY(p_pull_particles(h));
    const double *ps = h->host_ps.data();
    double *outs[] = {x, y, U, V, omega, Fx, Fy, torque};
    const int which[] = {PX_, PY_, PU_, PV_, POM_, PFX_, PFY_, PTQ_};
    for (int q = 0; q < 8; ++q) if (outs[q]) memcpy(outs[q], ps + (size_t)which[q] * N, (size_t)N * sizeof(double));
    return MGLC_OK;
}
extern "C" int mglc_p2d_set_forces(mglc_p2d *h, const double *Fx, const double *Fy, const double *torque) {
    if (!h || !Fx || !Fy || !torque) return MGLC_E_INVALID;
    const int N = h->p.N;
    MGLC_TRY(p_pull_particles(h));
    double *ps = h->host_ps.data();
    memcpy(ps + (size_t)PFX_ * N, Fx, (size_t)N * 8); memcpy(ps + (size_t)PFY_ * N, Fy, (size_t)N * 8); memcpy(ps + (size_t)PTQ_ * N, torque, (size_t)N * 8);
    return p_push_particles(h);
}

// ---- field transfers in the reference layout ------------------------------------------------------------------------------
static int p_lattice_io(Sub *S, double *host, double *dev, int rim, bool to_device) {
    if (!host) return MGLC_OK;
    const size_t bytes = (size_t)Q9 * (S->nx + 2 * rim) * (S->ny + 2 * rim) * sizeof(double);
    const dim3 grid((S->nx + 2 * rim + 127) / 128, S->ny + 2 * rim);
    if (to_device) {
        MGLC_CUDA(cudaMemcpyAsync(S->stage, host, bytes, cudaMemcpyHostToDevice, S->s));
        k_p_aos_to_soa<<<grid, 128, 0, S->s>>>(S->g, S->stage, rim, dev);
    } else {
        k_p_soa_to_aos<<<grid, 128, 0, S->s>>>(S->g, dev, rim, S->stage);
        MGLC_CUDA(cudaMemcpyAsync(host, S->stage, bytes, cudaMemcpyDeviceToHost, S->s));
    }
    S->launches += 1;
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
static int p_field_io(Sub *S, double *host, double *dev, bool to_device) {
    if (!host) return MGLC_OK;
    const size_t bytes = (size_t)S->nx * S->ny * sizeof(double);
    MGLC_CUDA(cudaMemcpyAsync(to_device ? (void *)dev : (void *)host, to_device ? (void *)host : (void *)dev, bytes,
                              to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
static int p_obst_io(Sub *S, int *host, int *dev, bool to_device) {
    if (!host) return MGLC_OK;
    const size_t bytes = (size_t)(S->nx + 2) * (S->ny + 2) * sizeof(int);
    const dim3 grid((S->nx + 2 + 127) / 128, S->ny + 2);
    if (to_device) MGLC_CUDA(cudaMemcpyAsync(S->istage, host, bytes, cudaMemcpyHostToDevice, S->s));
    k_p_obst_io<<<grid, 128, 0, S->s>>>(S->g, S->istage, dev, to_device ? 1 : 0);
    if (!to_device) MGLC_CUDA(cudaMemcpyAsync(host, S->istage, bytes, cudaMemcpyDeviceToHost, S->s));
    S->launches += 1;
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    return MGLC_OK;
}
extern "C" int mglc_p2d_upload(mglc_p2d *h, int r, const double *f, const double *f_post, const double *rho, const double *u,
                               const double *v, const int *obst) {
    Sub *S;
    MGLC_TRY(p_sub(h, r, &S));
    MGLC_TRY(p_lattice_io(S, const_cast<double *>(f), S->F, 3, true));
    MGLC_TRY(p_lattice_io(S, const_cast<double *>(f_post), S->Fp, 2, true));
    MGLC_TRY(p_field_io(S, const_cast<double *>(rho), S->rho, true));
    MGLC_TRY(p_field_io(S, const_cast<double *>(u), S->u, true));
    MGLC_TRY(p_field_io(S, const_cast<double *>(v), S->v, true));
    MGLC_TRY(p_obst_io(S, const_cast<int *>(obst), S->obst, true));
    return MGLC_OK;
}
extern "C" int mglc_p2d_download(mglc_p2d *h, int r, double *f, double *f_post, double *rho, double *u, double *v, int *obst) {
    Sub *S;
    MGLC_TRY(p_sub(h, r, &S));
    MGLC_TRY(p_lattice_io(S, f, S->F, 3, false));
    MGLC_TRY(p_lattice_io(S, f_post, S->Fp, 2, false));
    MGLC_TRY(p_field_io(S, rho, S->rho, false));
    MGLC_TRY(p_field_io(S, u, S->u, false));
    MGLC_TRY(p_field_io(S, v, S->v, false));
    MGLC_TRY(p_obst_io(S, obst, S->obst, false));
    return MGLC_OK;
}

// ---- the reference's subroutines --------------------------------------------------------------------------------------------
static inline dim3 grid_int(const Sub *S) { return dim3((S->nx + 127) / 128, S->ny); }
// the two summing kernels: at most ~8192 blocks (= partial sums), each walking several rows on a large lattice
static inline dim3 grid_rows(int nxt, int rows) {
    const int nbx = (nxt + 127) / 128;
    return dim3(nbx, std::max(1, std::min(rows, 8192 / nbx)));
}

extern "C" int mglc_p2d_initial(mglc_p2d *h) {
    if (!h) return MGLC_E_INVALID;
    // xCenterOld = xCenter, velocities 0, P4/initial.F90:84-115
    MGLC_TRY(p_pull_particles(h));
    const int N = h->p.N;
    double *ps = h->host_ps.data();
    for (int c = 0; c < N; ++c) {
        ps[PXO_ * N + c] = ps[PX_ * N + c]; ps[PYO_ * N + c] = ps[PY_ * N + c];
        const int zero[] = {PU_, PV_, POM_, PUO_, PVO_, POMO_, PFX_, PFY_, PTQ_, PSX_, PSY_, PST_};
        for (int z : zero) ps[(size_t)z * N + c] = 0.0;
    }
    MGLC_TRY(p_push_particles(h));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_initial<<<dim3((S->nx + 6 + 127) / 128, S->ny + 6), 128, 0, S->s>>>(S->g, h->p, S->bins, S->ps, S->F, S->Fp, S->obst, S->obstNew, S->rho,
                                                                              S->u, S->v, S->up, S->vp);
        MGLC_CUDA(cudaMemsetAsync(S->err, 0, sizeof(int), S->s));
        S->launches += 1;
    }
    return MGLC_OK;
}

extern "C" int mglc_p2d_collision(mglc_p2d *h) {
    if (!h) return MGLC_E_INVALID;
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_collision<<<grid_int(S), 128, 0, S->s>>>(S->g, h->p, S->F, S->rho, S->u, S->v, S->obst, S->Fp);
        S->launches += 1;
    }
    return MGLC_OK;
}

static int p_pack(Sub *S, cudaStream_t s) {
    for (int m = 0; m < 16; ++m) {
        const Msg &M = S->msgs[m];
        if (!M.send_count || M.skip) continue;
        const int *o = S->org[m];
        k_p_pack<<<(unsigned)((M.send_count + 255) / 256), 256, 0, s>>>(S->g, m < 8 ? S->Fp : S->F, o[0], o[1], o[4], o[5], M.sbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}
static int p_unpack(Sub *S, cudaStream_t s) {
    for (int m = 0; m < 16; ++m) {
        const Msg &M = S->msgs[m];
        if (!M.recv_count || M.skip) continue;
        const int *o = S->org[m];
        k_p_unpack<<<(unsigned)((M.recv_count + 255) / 256), 256, 0, s>>>(S->g, m < 8 ? S->Fp : S->F, o[2], o[3], o[4], o[5], M.rbuf);
        S->launches += 1;
    }
    return MGLC_OK;
}
// set = 0: send_all_fp (f_post, 2 deep), set = 1: send_all_f (f, 3 deep); P4/message_send_all.F90
static int p_exchange(mglc_p2d *h, int set) {
    if (h->nranks == 1) return MGLC_OK;
    P_EACH(h, S) for (int m = 0; m < 16; ++m) S->msgs[m].skip = (m / 8) != set;
    if (h->comm) {
        Sub *S = h->subs[0];
        MGLC_TRY(p_use(S));
        MGLC_TRY(p_pack(S, S->s));
        MGLC_TRY(halo_nccl_sendrecv(S->msgs, 16, h->comm, S->s));
        return p_unpack(S, S->s);
    }
    return halo_local_exchange(
        h->ports, [&](int r, cudaStream_t s) { return p_pack(h->subs[r], s); }, [&](int r, cudaStream_t s) { return p_unpack(h->subs[r], s); });
}
extern "C" int mglc_p2d_send_all_fp(mglc_p2d *h) { if (!h) return MGLC_E_INVALID; return p_exchange(h, 0); }
extern "C" int mglc_p2d_send_all_f(mglc_p2d *h) { if (!h) return MGLC_E_INVALID; return p_exchange(h, 1); }

// sum of rho over fluid nodes and their count on every subdomain, then across subdomains (2 x MPI_Allreduce)
static int p_fluid_average(mglc_p2d *h, bool use_new) {
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_fluid_sum<<<RED_BLOCKS, 256, 0, S->s>>>(S->g, S->rho, use_new ? S->obstNew : S->obst, S->part);
        k_p_fluid_sum_final<<<1, 1, 0, S->s>>>(RED_BLOCKS, S->part);
        S->launches += 2;
    }
    if (h->nranks == 1) return MGLC_OK;
    if (h->comm) {
        Sub *S = h->subs[0];
        MGLC_NCCL(ncclAllReduce(S->part, S->part, 2, ncclDouble, ncclSum, h->comm->nccl, S->s));
        return MGLC_OK;
    }
    double tot[2] = {0.0, 0.0};
    P_EACH(h, S) {                       // rank-ordered host sum, like the oracle
        MGLC_TRY(p_use(S));
        double e[2];
        MGLC_CUDA(cudaMemcpyAsync(e, S->part, sizeof e, cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        tot[0] += e[0]; tot[1] += e[1];
    }
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaMemcpyAsync(S->part, tot, sizeof tot, cudaMemcpyHostToDevice, S->s)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    return MGLC_OK;
}
extern "C" int mglc_p2d_set_rho_avg(mglc_p2d *h, double rhoAvg) {
    if (!h) return MGLC_E_INVALID;
    const double e[2] = {rhoAvg, 1.0};
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaMemcpyAsync(S->part, e, sizeof e, cudaMemcpyHostToDevice, S->s)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    return MGLC_OK;
}
extern "C" int mglc_p2d_get_rho_avg(mglc_p2d *h, double *rhoAvg) {
    if (!h || !rhoAvg) return MGLC_E_INVALID;
    Sub *S = h->subs[0];
    MGLC_TRY(p_use(S));
    double e[2];
    MGLC_CUDA(cudaMemcpyAsync(e, S->part, sizeof e, cudaMemcpyDeviceToHost, S->s));
    MGLC_CUDA(cudaStreamSynchronize(S->s));
    *rhoAvg = e[0] / e[1];
    return MGLC_OK;
}

template <int STAGES> static int p_update(mglc_p2d *h) {
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_update<STAGES><<<grid_int(S), 128, 0, S->s>>>(S->g, h->p, S->ps, S->ps, S->Fp, S->F, S->obst, S->rho, S->u, S->v, S->part, S->err, S->nlinks);
        S->launches += 1;
    }
    return MGLC_OK;
}
extern "C" int mglc_p2d_streaming(mglc_p2d *h) { if (!h) return MGLC_E_INVALID; return p_update<ST_STREAM>(h); }
extern "C" int mglc_p2d_bounceback(mglc_p2d *h) { if (!h) return MGLC_E_INVALID; return p_update<ST_WALLBB>(h); }
// recompute_rho_avg = 1 is the reference (rhoAvg from rho and obst, 2 Allreduce); 0 keeps the value set before (tests)
extern "C" int mglc_p2d_bounceback_particle(mglc_p2d *h, int recompute_rho_avg) {
    if (!h) return MGLC_E_INVALID;
    if (recompute_rho_avg) MGLC_TRY(p_fluid_average(h, false));
    return p_update<ST_PBB>(h);
}
extern "C" int mglc_p2d_macro(mglc_p2d *h) { if (!h) return MGLC_E_INVALID; return p_update<ST_MACRO>(h); }

static int p_zero_sums(mglc_p2d *h) {
    const int N = h->p.N;
    if (!N) return MGLC_OK;
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaMemsetAsync(S->ps + (size_t)PSX_ * N, 0, (size_t)3 * N * sizeof(double), S->s)); }
    return MGLC_OK;
}
// Allreduce(SUM) of the link sums (3 x cNumMax doubles in one message), then the per-particle tail
static int p_force_tail(mglc_p2d *h, int what) {
    const int N = h->p.N;
    if (!N) return MGLC_OK;
    if (h->nranks > 1 && (what & PT_FORCES)) {
        if (h->comm) {
            Sub *S = h->subs[0];
            MGLC_TRY(p_use(S));
            MGLC_NCCL(ncclAllReduce(S->ps + (size_t)PSX_ * N, S->ps + (size_t)PSX_ * N, (size_t)3 * N, ncclDouble, ncclSum, h->comm->nccl, S->s));
        } else {
            std::vector<double> tot((size_t)3 * N, 0.0), e((size_t)3 * N);
            P_EACH(h, S) {
                MGLC_TRY(p_use(S));
                MGLC_CUDA(cudaMemcpyAsync(e.data(), S->ps + (size_t)PSX_ * N, e.size() * 8, cudaMemcpyDeviceToHost, S->s));
                MGLC_CUDA(cudaStreamSynchronize(S->s));
                for (size_t q = 0; q < e.size(); ++q) tot[q] += e[q];
            }
            P_EACH(h, S) {
                MGLC_TRY(p_use(S));
                MGLC_CUDA(cudaMemcpyAsync(S->ps + (size_t)PSX_ * N, tot.data(), tot.size() * 8, cudaMemcpyHostToDevice, S->s));
                MGLC_CUDA(cudaStreamSynchronize(S->s));
            }
        }
    }
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        // two launches when both are asked for: the force pass reads every particle's position, the advance pass moves them
        for (int pass : {PT_FORCES, PT_ADVANCE}) {
            if (!(what & pass)) continue;
            if (pass == PT_FORCES && S->bins.B) k_p_forces_binned<<<(N + 63) / 64, 64, 0, S->s>>>(S->bins, h->p, S->ps, S->part, S->err);
            else k_p_particles<<<(N + 63) / 64, 64, 0, S->s>>>(h->p, S->ps, S->part, pass, S->err);
            S->launches += 1;
        }
        if (what & PT_ADVANCE) MGLC_TRY(p_fill_bins(h, S));      // the centres moved
    }
    return MGLC_OK;
}
extern "C" int mglc_p2d_calforce(mglc_p2d *h) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(p_zero_sums(h));
    MGLC_TRY(p_update<ST_FORCE>(h));
    return p_force_tail(h, PT_FORCES);
}
static int p_mask_refill(mglc_p2d *h) {
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_mask<<<dim3((S->nx + 2 + 127) / 128, S->ny + 2), 128, 0, S->s>>>(S->g, h->p, S->bins, S->ps, S->obstNew, S->rho, S->u, S->v);
        S->launches += 1;
    }
    MGLC_TRY(p_fluid_average(h, true));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_refill<<<grid_int(S), 128, 0, S->s>>>(S->g, h->p, S->bins, S->ps, S->obst, S->obstNew, S->F, S->rho, S->u, S->v, S->part, S->err, nullptr, nullptr);
        S->launches += 1;
        std::swap(S->obst, S->obstNew);      // obst = obstNew, P4/particle_update.F90:206 (obstNew is rebuilt from scratch next time)
    }
    return MGLC_OK;
}
extern "C" int mglc_p2d_update_center(mglc_p2d *h) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(p_force_tail(h, PT_ADVANCE));
    return p_mask_refill(h);
}

extern "C" int mglc_p2d_check(mglc_p2d *h, double *errorU) {
    if (!h || !errorU) return MGLC_E_INVALID;
    double t[2] = {0.0, 0.0};
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_check<<<RED_BLOCKS, 256, 0, S->s>>>(S->g, S->u, S->v, S->up, S->vp, S->obst, S->part + 1024);
        k_p_fluid_sum_final<<<1, 1, 0, S->s>>>(RED_BLOCKS, S->part + 1024);
        S->launches += 2;
        if (h->comm && h->nranks > 1) MGLC_NCCL(ncclAllReduce(S->part + 1024, S->part + 1024, 2, ncclDouble, ncclSum, h->comm->nccl, S->s));
        double e[2];
        MGLC_CUDA(cudaMemcpyAsync(e, S->part + 1024, sizeof e, cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        if (h->comm) { t[0] = e[0]; t[1] = e[1]; } else { t[0] += e[0]; t[1] += e[1]; }
    }
    *errorU = sqrt(t[0]) / sqrt(t[1]);
    return MGLC_OK;
}

// sum part[0..1] across subdomains (the two MPI_Allreduce of bounceback_particle / updateCenter)
static int p_reduce_part(mglc_p2d *h) {
    if (h->nranks == 1) return MGLC_OK;
    if (h->comm) {
        Sub *S = h->subs[0];
        MGLC_TRY(p_use(S));
        MGLC_NCCL(ncclAllReduce(S->part, S->part, 2, ncclDouble, ncclSum, h->comm->nccl, S->s));
        return MGLC_OK;
    }
    double tot[2] = {0.0, 0.0};
    P_EACH(h, S) {                       // rank-ordered host sum, like the oracle
        MGLC_TRY(p_use(S));
        double e[2];
        MGLC_CUDA(cudaMemcpyAsync(e, S->part, sizeof e, cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        tot[0] += e[0]; tot[1] += e[1];
    }
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaMemcpyAsync(S->part, tot, sizeof tot, cudaMemcpyHostToDevice, S->s)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    return MGLC_OK;
}

// One loop body of P4/main.F90:35-73, fused:
//   k_p_collision_sum   collision() + the rhoAvg sums of bounceback_particle()
//   [send_all_fp]       2-deep f_post halos
//   k_p_update<...>     streaming() + bounceback() + macro() per node, counting each node's links into solid nodes
//   k_p_links           bounceback_particle() + the link sums of calForce(), one block per particle, one thread per link
//   k_p_particles       Allreduced link sums -> spring forces, gravity, kinematics (calForce tail + updateCenter head)
//   [send_all_f]        3-deep f halos
//   k_p_mask_sum        new mask + the second rhoAvg sums;  k_p_refill   refill of uncovered nodes
// 6 launches on one GPU; every reduction has a fixed order, so a run is reproducible bit for bit.
static int p_enqueue_step(mglc_p2d *h) {
    const int N = h->p.N;
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_collision_sum<<<grid_rows(S->nx, S->ny), 128, 0, S->s>>>(S->g, h->p, S->F, S->rho, S->u, S->v, S->obst, S->Fp, S->gpart, S->ticket, S->part, S->rcount);
        S->launches += 1;
    }
    MGLC_TRY(p_exchange(h, 0));
    MGLC_TRY(p_reduce_part(h));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        static const bool generic_node_kernel = getenv("MGLC_P2D_NODE") && atoi(getenv("MGLC_P2D_NODE")) == 0;      // A/B runs
        if (generic_node_kernel)
            k_p_update<ST_STREAM | ST_WALLBB | ST_MACRO | ST_COUNT><<<grid_int(S), 128, 0, S->s>>>(
                S->g, h->p, S->ps, S->ps, S->Fp, S->F, S->obst, S->rho, S->u, S->v, S->part, S->err, S->nlinks);
        else k_p_node<<<grid_int(S), 128, 0, S->s>>>(S->g, h->p, S->Fp, S->F, S->obst, S->rho, S->u, S->v, S->nlinks);
        S->launches += 1;
        if (N) {
            k_p_links<<<N, LK_T, 0, S->s>>>(S->g, h->p, S->ps, S->ps, S->Fp, S->F, S->obst, S->nlinks, S->rho, S->u, S->v, S->part, S->err);
            S->launches += 1;
        }
    }
    if (N) {
        if (h->nranks > 1 || h->subs[0]->bins.B) MGLC_TRY(p_force_tail(h, PT_FORCES | PT_ADVANCE));      // Allreduce of the link sums + the two passes
        else {
            Sub *S = h->subs[0];
            k_p_particles_block<<<1, std::min(1024, 32 * N), 0, S->s>>>(h->p, S->ps, S->part, S->err);
            S->launches += 1;
        }
    }
    MGLC_TRY(p_exchange(h, 1));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_mask_sum<<<grid_rows(S->nx + 2, S->ny + 2), 128, 0, S->s>>>(S->g, h->p, S->bins, S->ps, S->obst, S->obstNew, S->rho, S->u, S->v,
                                                                      S->nlinks, S->err, S->gpart, S->ticket + 1, S->part, S->rlist, S->rcount, S->rcap);
        S->launches += 1;
    }
    MGLC_TRY(p_reduce_part(h));
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        k_p_refill<<<std::min(1184, (S->rcap + 127) / 128), 128, 0, S->s>>>(S->g, h->p, S->bins, S->ps, S->obst, S->obstNew, S->F, S->rho, S->u, S->v, S->part, S->err,
                                                                      S->rlist, S->rcount);
        S->launches += 1;
        std::swap(S->obst, S->obstNew);      // obst = obstNew, P4/particle_update.F90:206
    }
    return MGLC_OK;
}

// the previous fused schedule (one node kernel does the particle links too): kept for A/B runs, MGLC_P2D_LEGACY=1
static int p_step_legacy(mglc_p2d *h) {
    MGLC_TRY(mglc_p2d_collision(h));
    MGLC_TRY(p_exchange(h, 0));
    MGLC_TRY(p_fluid_average(h, false));
    MGLC_TRY(p_zero_sums(h));
    MGLC_TRY((p_update<ST_STREAM | ST_WALLBB | ST_PBB | ST_MACRO | ST_FORCE>(h)));
    MGLC_TRY(p_force_tail(h, PT_FORCES | PT_ADVANCE));
    MGLC_TRY(p_exchange(h, 1));
    return p_mask_refill(h);
}

// nsteps loop bodies.  A single-subdomain handle replays the step as a CUDA graph (one graph per obst/obstNew parity,
// captured the first time that parity comes up): the whole state is L2-resident and a step is a few microseconds of
// kernels, so launch overhead would otherwise dominate.
static int p_step_impl(mglc_p2d *h, int nsteps) {
    if (nsteps < 0) return MGLC_E_INVALID;
    static const bool legacy = getenv("MGLC_P2D_LEGACY") != nullptr, nograph = getenv("MGLC_P2D_NOGRAPH") != nullptr;
    // one subdomain per process: alone, or one rank of a communicator (the NCCL send/recv and all-reduces of the step are
    // captured with the kernels; every rank replays the same sequence)
    const bool graph = !legacy && !nograph && h->subs.size() == 1 && (h->nranks == 1 || h->comm) && !h->subs[0]->no_graph;
    for (int it = 0; it < nsteps; ++it) {
        if (legacy) { MGLC_TRY(p_step_legacy(h)); continue; }
        Sub *S = h->subs[0];
        if (!graph || S->no_graph) { MGLC_TRY(p_enqueue_step(h)); continue; }
        MGLC_TRY(p_use(S));
        const int par = S->obst == S->obst0 ? 0 : 1;
        if (!S->gexec[par]) {
            const long long l0 = S->launches;
            int *o0 = S->obst, *o1 = S->obstNew;
            cudaGraph_t gr = nullptr;
            MGLC_CUDA(cudaStreamBeginCapture(S->s, cudaStreamCaptureModeRelaxed));
            const int rc = p_enqueue_step(h);                 // swaps obst / obstNew like an executed step
            const cudaError_t ce = cudaStreamEndCapture(S->s, &gr);
            cudaError_t ie = cudaErrorUnknown;
            if (!rc && ce == cudaSuccess) ie = cudaGraphInstantiate(&S->gexec[par], gr, 0);
            if (gr) cudaGraphDestroy(gr);
            if (rc || ce != cudaSuccess || ie != cudaSuccess) {
                // nothing ran: put the host-side state back and run this handle without graphs from now on
                (void)cudaGetLastError();
                S->gexec[par] = nullptr;
                S->obst = o0; S->obstNew = o1; S->launches = l0;
                S->no_graph = 1;
                MGLC_TRY(p_enqueue_step(h));
                continue;
            }
            S->launches_per_step = (int)(S->launches - l0);
        } else {
            std::swap(S->obst, S->obstNew);
            S->launches += S->launches_per_step;
        }
        MGLC_CUDA(cudaGraphLaunch(S->gexec[par], S->s));
    }
    return MGLC_OK;
}
extern "C" int mglc_p2d_step(mglc_p2d *h, int nsteps) {
    if (!h) return MGLC_E_INVALID;
    MGLC_TRY(p_step_impl(h, nsteps));
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
extern "C" int mglc_p2d_step_timed(mglc_p2d *h, int nsteps, float *ms) {
    if (!h || !ms) return MGLC_E_INVALID;
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); }
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t0, S->s)); }
    MGLC_TRY(p_step_impl(h, nsteps));
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaEventRecord(S->ev_t1, S->s)); }
    float worst = 0.f;
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        MGLC_CUDA(cudaEventSynchronize(S->ev_t1));
        float t = 0.f;
        MGLC_CUDA(cudaEventElapsedTime(&t, S->ev_t0, S->ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return MGLC_OK;
}
// device-side fatal-physics flags (the reference prints and stops / MPI_Abort): 0 = none
extern "C" int mglc_p2d_error_flags(mglc_p2d *h, int *flags) {
    if (!h || !flags) return MGLC_E_INVALID;
    int all = 0;
    P_EACH(h, S) {
        MGLC_TRY(p_use(S));
        int e = 0;
        MGLC_CUDA(cudaMemcpyAsync(&e, S->err, sizeof e, cudaMemcpyDeviceToHost, S->s));
        MGLC_CUDA(cudaStreamSynchronize(S->s));
        all |= e;
    }
    *flags = all;
    if (all) { set_error("particle path: device error flags 0x%x (calQ=1 q=2 owner=4 interpenetration=8 wall=16 refill=32)", all); return MGLC_E_DIVERGED; }
    return MGLC_OK;
}
extern "C" int mglc_p2d_launch_count(mglc_p2d *h, long long *n) {
    if (!h || !n) return MGLC_E_INVALID;
    long long t = 0;
    P_EACH(h, S) t += S->launches;
    *n = t;
    return MGLC_OK;
}
extern "C" int mglc_p2d_sync(mglc_p2d *h) {
    if (!h) return MGLC_E_INVALID;
    P_EACH(h, S) { MGLC_TRY(p_use(S)); MGLC_CUDA(cudaStreamSynchronize(S->s)); MGLC_CUDA(cudaGetLastError()); }
    return MGLC_OK;
}
