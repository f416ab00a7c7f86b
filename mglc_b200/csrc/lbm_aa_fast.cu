// lbm_aa_fast.cu -- throughput build (-fmad=true) of the AA-pattern D3Q19 kernels; see lbm_aa_kernels.inl
#define MGLC_NS fast
#include "lbm_aa_kernels.inl"
