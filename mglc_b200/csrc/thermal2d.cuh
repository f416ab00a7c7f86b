// thermal2d.cuh -- internal: parameters and launcher prototypes of the 2-D thermal D2Q9 + D2Q5 path (thermal2d.cu,
// thermal2d_fast.cu).  B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/.  Geometry is lid2d.cuh's Geom2.
#pragma once
#include "lid2d.cuh"

namespace mglc {

struct T2Params {
    double Snu, Sq, Qd, Qnu, paraA, gBeta, Tref, rho0, Thot, Tcold;   // module.F90:67-81
    double wallT[4];     // (4+paraA)/10 * T_wall per side (+x, -x, +y, -y), evolution_g.F90:99-140; used when bcT[side] != 0
    int bcT[4];          // MGLC_BCT_ADIABATIC or a constant-temperature kind (0 for a periodic side)
    int perx;            // 1 = vertical walls periodic for f and g (seq/bouyancy2d_acc.F90:777-791, 1037-1045); the subdomain spans x
    int variant;         // MGLC_T2D_MPI | MGLC_T2D_ACC (collision() rounds f_post(0) term by term, acc:679)
    // the sheared Rayleigh-Benard programs (RB2 = seq/R_B_2d.F90, also seq/bouyancy2d_omp.F90): walls that move along themselves.
    // Uwall = TopLeft, TopRight, BottomLeft, BottomRight (u of the horizontal walls; left half: global i <= nxHalf),
    // LeftTop, LeftBottom, RightTop, RightBottom (v of the vertical walls; bottom half: global j <= nyHalf)   RB2:87,118-120
    double Uwall[8];
    int moving;          // any Uwall != 0: diagonal populations off a wall get - rho*C/6 (RB2:790-898), rho of the previous macro()
    int cornersT;        // RB2:1086-1106: in a corner cell the population off the vertical wall takes the plate's constant-T rule
    int start[2], total[2];   // this subdomain's 0-based offset in the global lattice and the global size (per subdomain copy)
};
// velocity of the horizontal wall (top / bottom) at global column gi, of the vertical wall (right / left) at global row gj
__host__ __device__ inline double t2_wall_u(const T2Params &p, bool top, int gi) {
    const bool left = gi <= (p.total[0] - 1) / 2 + 1;
    return p.Uwall[top ? (left ? 0 : 1) : (left ? 2 : 3)];
}
__host__ __device__ inline double t2_wall_v(const T2Params &p, bool right, int gj) {
    const bool bottom = gj <= (p.total[1] - 1) / 2 + 1;
    return p.Uwall[right ? (bottom ? 7 : 6) : (bottom ? 5 : 4)];
}
// C of "f(a) = f_post(opp) - rho*C/6" for a diagonal population (ex, ey) whose upstream cell lies beyond a vertical wall (hx),
// a horizontal wall (hy) or both (the corner): RB2:795-814 (vertical), :840-859 (horizontal), :870-896 (corners)
__host__ __device__ inline double t2_wall_coef(const T2Params &p, int ex, int ey, bool hx, bool hy, int gi, int gj) {
    const double cu = -((double)ex * t2_wall_u(p, ey == -1, gi)), cv = -((double)ey * t2_wall_v(p, ex == -1, gj));
    return (hx && hy) ? cu + cv : hy ? cu : cv;
}

// Fy is updated in place by the fused kernel (a cell reads and writes only its own entry); Fx is identically 0 after any
// collision() (evolution_f.F90:45) and is only materialised by the un-fused collision kernel.
#define MGLC_DECLARE_T2_LAUNCHERS                                                                                              \
    int launch_t2_collision(const Geom2 &g, const T2Params &p, const double *F, const double *rho, const double *u,           \
                            const double *v, const double *T, double *Fpost, double *Fx, double *Fy, cudaStream_t s);         \
    int launch_t2_collisionT(const Geom2 &g, const T2Params &p, const double *G, const double *u, const double *v,            \
                             const double *T, double *Gpost, cudaStream_t s);                                                 \
    /* streaming+bounceback+streamingT+bouncebackT+macro+macroT of step n, collision+collisionT of step n+1 */                \
    /* rho: the field array; with moving walls its wall cells are read (previous macro()) and rewritten every launch */     \
    int launch_t2_fused(const Geom2 &g, const T2Params &p, const double *Fin, double *Fout, const double *Gin, double *Gout,  \
                        double *Fy, double *rho, cudaStream_t s);                                                             \
    /* epilogue of a fused run: the same pulls + macro + macroT -> F, G (pre-collision) and the fields */                     \
    int launch_t2_stream_macro(const Geom2 &g, const T2Params &p, const double *Fin, double *F, const double *Gin, double *G, \
                               const double *Fy, double *rho, double *u, double *v, double *T, cudaStream_t s);
namespace strict { MGLC_DECLARE_T2_LAUNCHERS }
namespace fast { MGLC_DECLARE_T2_LAUNCHERS }

}  // namespace mglc
