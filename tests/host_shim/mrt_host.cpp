// Test-only shim: compiles the product's per-cell D3Q19 arithmetic (mglc_b200/csrc/d3q19_mrt.inl) for
// the HOST so the CPU-only test suite can check it against the oracle without a GPU.  Two copies of the
// same source: MGLC_STRICT (reference operation order; must be bit-identical to the oracle when built
// with -ffp-contract=off) and the fast restructuring (must agree to rounding).
#define __device__
#define __forceinline__ inline
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

namespace strict_ns {
#define MGLC_STRICT 1
#include "../../mglc_b200/csrc/d3q19_mrt.inl"
#undef MGLC_STRICT
}
namespace fast_ns {
#include "../../mglc_b200/csrc/d3q19_mrt.inl"
}

extern "C" {
void shim_collide(int strict, const double *f, double rho, double u, double v, double w, double Snu, double Sq,
                  double *fp) {
    double fi[19], fo[19];
    for (int a = 0; a < 19; ++a) fi[a] = f[a];
    if (strict) strict_ns::d3q19_collide(fi, rho, u, v, w, Snu, Sq, fo);
    else fast_ns::d3q19_collide(fi, rho, u, v, w, Snu, Sq, fo);
    for (int a = 0; a < 19; ++a) fp[a] = fo[a];
}
void shim_collide_bgk(int strict, const double *f, double rho, double u, double v, double w, double Snu, double *fp) {
    double fi[19], fo[19];
    for (int a = 0; a < 19; ++a) fi[a] = f[a];
    if (strict) strict_ns::d3q19_collide_bgk(fi, rho, u, v, w, Snu, fo);
    else fast_ns::d3q19_collide_bgk(fi, rho, u, v, w, Snu, fo);
    for (int a = 0; a < 19; ++a) fp[a] = fo[a];
}
void shim_macro(const double *f, double *out4) {
    double fi[19];
    for (int a = 0; a < 19; ++a) fi[a] = f[a];
    strict_ns::d3q19_macro(fi, out4[0], out4[1], out4[2], out4[3]);
}
}
