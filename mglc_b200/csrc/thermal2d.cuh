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
};

// Fy is updated in place by the fused kernel (a cell reads and writes only its own entry); Fx is identically 0 after any
// collision() (evolution_f.F90:45) and is only materialised by the un-fused collision kernel.
#define MGLC_DECLARE_T2_LAUNCHERS                                                                                              \
    int launch_t2_collision(const Geom2 &g, const T2Params &p, const double *F, const double *rho, const double *u,           \
                            const double *v, const double *T, double *Fpost, double *Fx, double *Fy, cudaStream_t s);         \
    int launch_t2_collisionT(const Geom2 &g, const T2Params &p, const double *G, const double *u, const double *v,            \
                             const double *T, double *Gpost, cudaStream_t s);                                                 \
    /* streaming+bounceback+streamingT+bouncebackT+macro+macroT of step n, collision+collisionT of step n+1 */                \
    int launch_t2_fused(const Geom2 &g, const T2Params &p, const double *Fin, double *Fout, const double *Gin, double *Gout,  \
                        double *Fy, cudaStream_t s);                                                                          \
    /* epilogue of a fused run: the same pulls + macro + macroT -> F, G (pre-collision) and the fields */                     \
    int launch_t2_stream_macro(const Geom2 &g, const T2Params &p, const double *Fin, double *F, const double *Gin, double *G, \
                               const double *Fy, double *rho, double *u, double *v, double *T, cudaStream_t s);
namespace strict { MGLC_DECLARE_T2_LAUNCHERS }
namespace fast { MGLC_DECLARE_T2_LAUNCHERS }

}  // namespace mglc
