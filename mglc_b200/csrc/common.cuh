// common.cuh -- internal: device geometry of one subdomain, error plumbing, kernel launch prototypes.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include "../../include/mglc.h"

namespace mglc {

constexpr int Q = 19;
// x-index of interior cell i=1 inside a padded row: rows start 128-byte aligned and so does i=1, so
// every warp-wide store of 32 consecutive cells is line-aligned.  Halo i=0 sits at OX-1.
constexpr int OX = 16;

// SoA population layout F[a][k][j][x], one-cell halo in every dimension (k,j in 0..n+1), x padded.
struct Geom {
    int nx, ny, nz;        // interior size
    int px;                // x pitch in doubles (multiple of 16)
    int py, pz;            // ny+2, nz+2
    long long sy, sz, sq;  // strides in doubles: row, plane, population
    // wall flags: 1 if the face is a physical wall of the global box (coords==0 / dims-1), L3/bounce_back.f90
    int wall[6];           // +x,-x,+y,-y,+z,-z  (same order as nbr_surface)
    int lid;               // 1 if this subdomain holds the moving lid (coords(2)==dims(2)-1)
    __host__ __device__ long long idx(int a, int i, int j, int k) const {
        return a * sq + k * sz + j * sy + (i + OX - 1);
    }
    __host__ __device__ long long cell(int i, int j, int k) const {   // macro fields (nx,ny,nz), 1-based
        return (long long)(i - 1) + (long long)nx * ((j - 1) + (long long)ny * (k - 1));
    }
};

inline Geom make_geom(int nx, int ny, int nz) {
    Geom g{};
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.px = ((nx + OX + 1 + 15) / 16) * 16;
    g.py = ny + 2; g.pz = nz + 2;
    g.sy = g.px; g.sz = (long long)g.px * g.py; g.sq = g.sz * g.pz;
    return g;
}

// Direct halo stores (fused step on several GPUs): where a neighbouring subdomain's lattice is mapped into this
// GPU's address space (peer access inside one process, CUDA IPC between processes, both over NVLink), the fused
// kernel writes the outgoing populations of its boundary cells straight into the neighbour's halo cells instead
// of leaving them to pack -> ncclSend/ncclRecv -> unpack.  Entry d = direction of the message: 0..5 faces
// +x,-x,+y,-y,+z,-z (5 populations, ex_sendrecv.f90:12-59; thermal: + g population d+1, B3:1421-1468),
// 7..18 the edge population d crosses (ex_sendrecv.f90:64-123).  The neighbour's block can differ in size by one
// cell per axis (decompose_1d), so its strides and extents are carried too.
struct PeerTable {
    unsigned mask;             // bit d set: there is a neighbour in direction d
    double *F[19];             // the neighbour's f_post lattice being written this step (same ping-pong parity)
    double *G[6];              // thermal: its g_post lattice
    long long sy[19], sz[19], sq[19];
    int n[19][3];              // the neighbour's interior size
    // halo push (transport 3) only: where the two x-face messages go instead of the neighbour's halo column.  A halo column is
    // one double every `sy` doubles, and 8-byte remote stores scattered like that cost ~2 ns each over NVLink (2.9 M of them per
    // x face at 768^3: 5.6 ms, against 0.01 ms for a contiguous y or z face, profiles/r2g_*).  So the x faces are written
    // contiguously -- [slot][k][j], the layout of the NCCL messages -- into a staging area in the neighbour's memory, and the
    // neighbour scatters them into its own halo column after the barrier (local stores).  XS[d] = staging of face message d
    // (0: +x, 1: -x) for the lattice this table belongs to, nullptr = store into the halo column as usual.
    double *XS[2];
    // sticky error word of the neighbour barrier (k_halo_wait): once set, a launch stores nothing into a neighbour any
    // more (it may still be reading those halos); the subdomain's own numbers are meaningless from then on and the host
    // reports MGLC_E_STATE at every point where it synchronises, so such a run never comes back as MGLC_OK
    const int *err;
};

// the flag words of the neighbour barrier that goes with PeerTable: signal[d] = my slot in the memory of the
// neighbour in direction d, wait[d] = the slot that neighbour raises in mine
struct SyncTable {
    unsigned mask;
    unsigned long long *signal[19];
    unsigned long long *wait[19];
};

struct LbmParams {
    double Snu, Sq, U0, rho0;
    int bgk;             // 1 = MGLC_BGK (L3/collision.f90:191-198), 0 = the MRT operator
};

// thermal double-distribution path (MGLC_D3Q19_D3Q7), B3 = MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90
constexpr int QT = 7;
struct ThermalParams {
    double Snu, Sq, Qd, Qnu, paraA, gBeta, Tref, omegaRot;   // B3:30-47,73-74
    double Thot, Tcold;
    double wallT[6];     // (6+paraA)/21 * T_wall per face (+x,-x,+y,-y,+z,-z), B3:1128-1163; used when bcT[face] != 0
    int bcT[6];          // MGLC_BCT_ADIABATIC or a constant-temperature kind
};

// D3Q19 tables, L3/commondata.f90:32-40 (host copies; device code uses constexpr switch tables)
static const int h_ex[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int h_ey[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int h_ez[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int h_opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
// populations crossing each face, ascending = tag order of L3/ex_sendrecv.f90:12,20,29,37,46,54
static const int h_face_pops[6][5] = {{1, 7, 9, 11, 13}, {2, 8, 10, 12, 14}, {3, 7, 8, 15, 17},
                                      {4, 9, 10, 16, 18}, {5, 11, 12, 15, 16}, {6, 13, 14, 17, 18}};

void set_error(const char *fmt, ...);

#define MGLC_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ::mglc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorMemoryAllocation) ? MGLC_E_NOMEM : MGLC_E_CUDA;            \
        }                                                                                     \
    } while (0)

// ---- kernel launchers (each returns the number of kernels it launched) ----
// Two builds of the same kernel source: namespace strict (-fmad=false, true divisions: bit-identical
// to the CPU oracle) and namespace fast (reciprocal multiplies + FMA).
#define MGLC_DECLARE_LBM_LAUNCHERS                                                                   \
    int launch_collision(const Geom &g, const LbmParams &p, const double *F, const double *rho,      \
                         const double *u, const double *v, const double *w, double *Fpost,           \
                         cudaStream_t s);                                                            \
    /* stream + macro + collide: pull from Fin (halo'd, post-collision) -> post-collision Fout;      \
       box = [i0,i1]x[j0,j1]x[k0,k1] inclusive, 1-based interior cells */                            \
    int launch_fused(const Geom &g, const LbmParams &p, const double *Fin, double *Fout,             \
                     const double *rho_lid_in, double *rho_lid_out, const int box[6], cudaStream_t s, \
                     const PeerTable *peers = nullptr);   /* device pointer; non-null: direct halo stores */ \
    /* stream + macro (epilogue of a fused run): Fin (post-collision) -> F (pre-collision) + fields */\
    int launch_stream_macro(const Geom &g, const LbmParams &p, const double *Fin, double *F,         \
                            const double *rho_lid_in, double *rho, double *u, double *v, double *w,  \
                            cudaStream_t s);

// thermal launchers: Fc = the three force fields Fx,Fy,Fz stored back to back (ncell doubles each)
#define MGLC_DECLARE_THERMAL_LAUNCHERS                                                               \
    int launch_th_collision(const Geom &g, const ThermalParams &tp, const double *F, const double *rho, \
                            const double *u, const double *v, const double *w, const double *T,     \
                            double *Fpost, double *Fc, cudaStream_t s);                              \
    int launch_th_collisionT(const Geom &g, const ThermalParams &tp, const double *G, const double *u, \
                             const double *v, const double *w, const double *T, double *Gpost,       \
                             cudaStream_t s);                                                        \
    /* streaming+bounceback+streamingT+bouncebackT+macro+macroT of step n, collision+collisionT of n+1 */ \
    int launch_th_fused(const Geom &g, const ThermalParams &tp, const double *Fin, double *Fout,     \
                        const double *Gin, double *Gout, const double *Fc_in, double *Fc_out,        \
                        const int box[6], cudaStream_t s, const PeerTable *peers = nullptr);         \
    int launch_th_stream_macro(const Geom &g, const ThermalParams &tp, const double *Fin, double *F, \
                               const double *Gin, double *G, const double *Fc_in, double *rho,       \
                               double *u, double *v, double *w, double *T, cudaStream_t s);

namespace strict { MGLC_DECLARE_LBM_LAUNCHERS MGLC_DECLARE_THERMAL_LAUNCHERS }
namespace fast { MGLC_DECLARE_LBM_LAUNCHERS MGLC_DECLARE_THERMAL_LAUNCHERS }

// exact (copy / order-preserving) kernels, built once with -fmad=false
int launch_initial(const Geom &g, const LbmParams &p, double *F, double *rho, double *u, double *v,
                   double *w, cudaStream_t s);
int launch_streaming(const Geom &g, const double *Fpost, double *F, cudaStream_t s);
int launch_bounceback(const Geom &g, const LbmParams &p, const double *Fpost, const double *rho, double *F,
                      cudaStream_t s);
int launch_macro(const Geom &g, const double *F, double *rho, double *u, double *v, double *w, cudaStream_t s);
// check(): partial[0] += sum (du^2+dv^2), partial[1] += sum(u^2+v^2+w^2); then up<-u, vp<-v, wp<-w
int launch_check(const Geom &g, const double *u, const double *v, const double *w, double *up, double *vp,
                 double *wp, double *partial2, cudaStream_t s);
// halo pack/unpack for one message (dir 0..5 faces, 7..18 edges), buffer layout [slot][t2][t1]
int launch_pack(const Geom &g, const double *Fpost, int dir, double *buf, cudaStream_t s);
int launch_unpack(const Geom &g, double *Fpost, int dir, const double *buf, cudaStream_t s);
// AoS (reference layout) <-> SoA transposes over a linear cell range [c0, c0+ncells); nq = 19 or 7
int launch_aos_to_soa(const Geom &g, int nq, const double *aos_chunk, double *F, long long c0, long long ncells,
                      int with_halo, cudaStream_t s);
int launch_soa_to_aos(const Geom &g, int nq, const double *F, double *aos_chunk, long long c0, long long ncells,
                      int with_halo, cudaStream_t s);
// thermal exact kernels (B3): initial(), streamingT(), bouncebackT(), macro() with F/2, macroT(), check()
int launch_th_initial(const Geom &g, const ThermalParams &tp, double *F, double *G, double *rho, double *u, double *v,
                      double *w, double *T, cudaStream_t s);
int launch_streamingT(const Geom &g, const double *Gpost, double *G, cudaStream_t s);
int launch_bouncebackT(const Geom &g, const ThermalParams &tp, const double *Gpost, double *G, cudaStream_t s);
int launch_th_macro(const Geom &g, const double *F, const double *Fc, double *rho, double *u, double *v, double *w,
                    cudaStream_t s);
int launch_macroT(const Geom &g, const double *G, double *T, cudaStream_t s);
// partial[0..3] = sum |du|^2 (u,v,w), sum |u|^2, sum |dT|, sum |T|; then up,vp,wp,Tp <- u,v,w,T
int launch_th_check(const Geom &g, const double *u, const double *v, const double *w, const double *T, double *up,
                    double *vp, double *wp, double *Tp, double *partial4, cudaStream_t s);
// g halo messages: face dir 0..5 carries the single population dir+1 (B3:1421-1468)
int launch_pack_g(const Geom &g, const double *Gpost, int face, double *buf, cudaStream_t s);
int launch_unpack_g(const Geom &g, double *Gpost, int face, const double *buf, cudaStream_t s);
int launch_nure(const Geom &g, const double *u, const double *v, const double *w, const double *T, double *part, cudaStream_t s);
// one launch that packs (or unpacks) every message of an exchange: dir 0..18 = f messages, 20 + face = g messages
struct MsgBatch { int n; int dir[24]; double *buf[24]; };
int launch_pack_all(const Geom &g, const MsgBatch &mb, double *Fpost, double *Gpost, bool unpack, cudaStream_t s);
int launch_push_halos(const Geom &g, const PeerTable *pt_dev, const double *F, const double *G, cudaStream_t s);
int launch_fill(double *p, long long n, double value, cudaStream_t s);
int launch_halo_signal(const SyncTable &t, unsigned long long epoch, cudaStream_t s);
int launch_halo_wait(const SyncTable &t, unsigned long long epoch, int *err, cudaStream_t s);
double halo_timeout_seconds();
int check_scratch_doubles();
void msg_dims(const Geom &g, int dir, int &n1, int &n2, int &npop);

}  // namespace mglc
