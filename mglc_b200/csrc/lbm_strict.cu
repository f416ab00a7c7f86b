// lbm_strict.cu -- the lattice-update kernels with the reference's exact operation order
// (build: -fmad=false; divisions stay IEEE divisions).  Bit-identical to the CPU oracle.
#define MGLC_NS strict
#define MGLC_STRICT 1
#include "lbm_kernels.inl"
