// lid2d_fast.cu -- the 2-D D2Q9 collision / fused kernels restructured for throughput (build: -fmad=true, constant
// reciprocals, shared partial sums).  Fields agree with the strict build to <= 1e-12 rel. L2.
#define MGLC_NS fast
#include "lid2d_kernels.inl"
