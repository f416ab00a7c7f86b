// lbm_fast.cu -- the lattice-update kernels restructured for throughput (build: -fmad=true,
// constant reciprocals, shared partial sums).  Fields agree with the strict build to <= 1e-12 rel. L2.
#define MGLC_NS fast
#include "lbm_kernels.inl"
