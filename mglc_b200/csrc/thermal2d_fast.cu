// thermal2d_fast.cu -- throughput build (-fmad=true) of the 2-D thermal kernels; see thermal2d_kernels.inl
#define MGLC_NS fast
#include "thermal2d_kernels.inl"
