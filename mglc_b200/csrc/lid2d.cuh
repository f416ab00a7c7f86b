// lid2d.cuh -- internal: geometry and launcher prototypes of the 2-D D2Q9 lid-driven cavity path (lid2d.cu, lid2d_fast.cu).
#pragma once
#include "common.cuh"

namespace mglc {

// SoA population layout F[a][j][x]: j in 0..ny+1 (one-cell halo ring), x padded as in the 3-D lattice (OX, common.cuh)
struct Geom2 {
    int nx, ny;
    int px;                 // x pitch in doubles (multiple of 16)
    long long sy, sq;       // row and population strides in doubles
    int wall[4];            // +x, -x, +y (the lid), -y : 1 if that side is a physical wall of the global box
    __host__ __device__ long long idx(int a, int i, int j) const { return a * sq + j * sy + (i + OX - 1); }
    __host__ __device__ long long cell(int i, int j) const { return (long long)(i - 1) + (long long)nx * (j - 1); }   // rho,u,v (nx,ny)
};
inline Geom2 make_geom2(int nx, int ny) {
    Geom2 g{};
    g.nx = nx; g.ny = ny;
    g.px = ((nx + OX + 1 + 15) / 16) * 16;
    g.sy = g.px; g.sq = (long long)g.px * (ny + 2);
    return g;
}
struct L2Params { double Snu, Sq, U0, rho0; };

#define MGLC_DECLARE_L2_LAUNCHERS                                                                                              \
    int launch_l2_collision(const Geom2 &g, const L2Params &p, int variant, const double *F, const double *rho, const double *u, \
                            const double *v, double *Fpost, cudaStream_t s);                                                   \
    int launch_l2_fused(const Geom2 &g, const L2Params &p, int variant, const double *Fin, double *Fout, const double *lid_in,   \
                        double *lid_out, cudaStream_t s);                                                                      \
    int launch_l2_stream_macro(const Geom2 &g, const L2Params &p, int variant, const double *Fin, double *F,                  \
                               const double *lid_in, double *rho, double *u, double *v, cudaStream_t s);
namespace strict { MGLC_DECLARE_L2_LAUNCHERS }
namespace fast { MGLC_DECLARE_L2_LAUNCHERS }

}  // namespace mglc
