// Test-only shim: the product's calQ (mglc_b200/csrc/p2d_calq.inl: bisection with square-root-free far-field
// halvings) compiled for the HOST, so the CPU-only suite can check it against the oracle's verbatim restatement
// of P4/particle_bounceback.F90:98-141 bit for bit.
#include <cmath>
#define __device__
enum { ERR_CALQ = 1, ERR_Q = 2 };
#include "../../mglc_b200/csrc/p2d_calq.inl"
extern "C" int shim_calq(double xc, double yc, double rad, double i, double j, double exa, double eya, double *out3) {
    double x0 = 0, y0 = 0, q = 0;
    const int rc = calQ_link(xc, yc, rad, i, j, exa, eya, x0, y0, q);
    out3[0] = x0; out3[1] = y0; out3[2] = q;
    return rc;
}
