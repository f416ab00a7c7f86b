// lbm_aa.cuh -- internal: launcher prototypes of the AA-pattern (single-lattice) D3Q19 path (lbm_aa.cu, lbm_aa_fast.cu)
#pragma once
#include "common.cuh"

namespace mglc {
#define MGLC_DECLARE_AA_LAUNCHERS                                                                                              \
    /* peers (a PeerTable on the HOST, passed by value) != nullptr: a decomposed lattice, neighbours mapped */               \
    int launch_aa_collide0(const Geom &g, const LbmParams &p, double *A, const double *rho, const double *u, const double *v, \
                           const double *w, cudaStream_t s, const PeerTable *peers = nullptr);                                \
    int launch_aa_even(const Geom &g, const LbmParams &p, double *A, double *rho_lid_out, cudaStream_t s,                     \
                       const PeerTable *peers = nullptr);                                                                     \
    int launch_aa_odd(const Geom &g, const LbmParams &p, double *A, const double *rho_lid_in, cudaStream_t s,                 \
                      const PeerTable *peers = nullptr);
namespace strict { MGLC_DECLARE_AA_LAUNCHERS }
namespace fast { MGLC_DECLARE_AA_LAUNCHERS }

// The launch schedule of nsteps loop bodies (L3/main.f90:89-97) on the single lattice; shared by lbm_aa.cu and by the CPU
// emulation in tests/host_shim/aa_host.cpp.
//   AA_NATURAL  A[a][x] = f_a(x) of loop body n, fields of body n
//   AA_POST     A[opp(a)][x] = f_post_a(x) of body n (collision done, not streamed), fields of body n, lid = rho plane of body n-1
// From NATURAL: the lid plane is taken from the rho field and collision() runs with the stored fields; then launches alternate
// odd (POST -> NATURAL: one collision, two streaming steps) and even (NATURAL -> POST: one collision).  From POST the run
// resumes with an odd launch.  A run that ends in NATURAL finishes with macro() on the lattice; one that ends in POST
// computes the fields through streaming() + bounceback() + macro() and leaves the lattice as it is.
enum { AA_NATURAL = 0, AA_POST = 1 };
enum AaOp { AA_OP_LID_PLANE, AA_OP_COLLIDE0, AA_OP_ODD, AA_OP_EVEN, AA_OP_MACRO_POST, AA_OP_MACRO };
template <class Launch>
inline long long aa_run(int &layout, int nsteps, Launch op) {
    long long launches = 0;
    if (nsteps <= 0) return 0;
    int collisions = 0;
    if (layout == AA_NATURAL) {
        launches += op(AA_OP_LID_PLANE);
        launches += op(AA_OP_COLLIDE0);
        layout = AA_POST;
        collisions = 1;
    }
    while (collisions < nsteps) {
        if (layout == AA_POST) { launches += op(AA_OP_ODD); layout = AA_NATURAL; }
        else { launches += op(AA_OP_EVEN); layout = AA_POST; }
        ++collisions;
    }
    launches += op(layout == AA_POST ? AA_OP_MACRO_POST : AA_OP_MACRO);
    return launches;
}
}  // namespace mglc
