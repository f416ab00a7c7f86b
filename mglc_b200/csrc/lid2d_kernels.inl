// lid2d_kernels.inl -- 2-D D2Q9 MRT lid-driven cavity kernels; compiled twice (lid2d.cu: namespace strict, -fmad=false;
// lid2d_fast.cu: namespace fast, -fmad=true).
//   L2C = MPI/Lid_driven_cavity/c/lid_driven_cavity.c            L2F = MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/
//   L2I = MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90 (the incompressible model, sequential)
// Device layout: SoA F[a][j][x], one-cell halo ring, rows padded like the 3-D lattice (interior cell i = 1 at x-index OX,
// pitch a multiple of 16 doubles): thread <-> cell, threadIdx.x along x, every warp store 128-byte aligned.
// The fused kernel is the reference loop body rotated by half a step (like k_fused of the 3-D path): streaming() + bounceback()
// / boundary() + macro() of step n and collision() of step n+1 in one pass: 9 loads + 9 stores of fp64 = 144 B per cell.
#include "lid2d.cuh"

namespace mglc {
namespace MGLC_NS {

// collision() of one cell.  VARIANT 0: L2C c:186-255 (meq(8) = u*v, the inverse as (...)/36.0);
// VARIANT 1: L2F evolution.f90:16-66 (grouped sums, per-term divisions, meq(8) = rho*(u*v));
// VARIANT 2: L2I:184-233 (the sums and divisions of L2F, the equilibrium without the rho factors, L2I:195-202);
// VARIANT 3: L2C with its model switch set to SRT (c:13-14), c:160-176: f_post = f - 1.0/tau*(f - feq), 1.0/tau = Snu (c:98).
template <int VARIANT>
__device__ __forceinline__ void d2q9_collide(const double (&f)[9], double rho, double u, double v, double Snu, double Sq,
                                             double (&fp)[9]) {
    if (VARIANT == 3) {
        const double omega[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
        const double ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};      // c:25-26 (doubles there too)
#ifdef MGLC_STRICT
        const double u2 = u * u + v * v;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const double ue = u * ex[a] + v * ey[a];
            const double feq = rho * omega[a] * (1.0 + 3.0 * ue + 4.5 * ue * ue - 1.5 * u2);
            fp[a] = f[a] - Snu * (f[a] - feq);
        }
#else
        const double base = 1.0 - 1.5 * (u * u + v * v);
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const double ue = u * ex[a] + v * ey[a];
            const double feq = (rho * omega[a]) * (base + ue * (3.0 + 4.5 * ue));
            fp[a] = f[a] + Snu * (feq - f[a]);
        }
#endif
        return;
    }
#ifdef MGLC_STRICT
    double m[9], meq[9], mp[9];
    const double s[9] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, Snu, Snu};
    if (VARIANT == 0) {
        m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        m[1] = -4 * f[0] - f[1] - f[2] - f[3] - f[4] + 2 * f[5] + 2 * f[6] + 2 * f[7] + 2 * f[8];
        m[2] = 4 * f[0] - 2 * f[1] - 2 * f[2] - 2 * f[3] - 2 * f[4] + f[5] + f[6] + f[7] + f[8];
    } else {
        m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * (f[5] + f[6] + f[7] + f[8]);
        m[2] = 4.0 * f[0] - 2.0 * (f[1] + f[2] + f[3] + f[4]) + f[5] + f[6] + f[7] + f[8];
    }
    m[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
    m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
    m[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
    m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
    m[7] = f[1] - f[2] + f[3] - f[4];
    m[8] = f[5] - f[6] + f[7] - f[8];
    meq[0] = rho;
    if (VARIANT == 2) {
        meq[1] = -2.0 * rho + 3.0 * (u * u + v * v);
        meq[2] = rho - 3.0 * (u * u + v * v);
        meq[3] = u;
        meq[4] = -u;
        meq[5] = v;
        meq[6] = -v;
        meq[7] = u * u - v * v;
        meq[8] = u * v;
    } else {
        meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
        meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
        meq[3] = rho * u;
        meq[4] = -(rho * u);
        meq[5] = rho * v;
        meq[6] = -(rho * v);
        meq[7] = rho * (u * u - v * v);
        meq[8] = VARIANT == 0 ? u * v : rho * (u * v);
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);
    if (VARIANT == 0) {
        fp[0] = (mp[0] - mp[1] + mp[2]) / 9.0;
        fp[1] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] + 6.0 * mp[3] - 6.0 * mp[4] + 9.0 * mp[7]) / 36.0;
        fp[2] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] + 6.0 * mp[5] - 6.0 * mp[6] - 9.0 * mp[7]) / 36.0;
        fp[3] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] - 6.0 * mp[3] + 6.0 * mp[4] + 9.0 * mp[7]) / 36.0;
        fp[4] = (4.0 * mp[0] - mp[1] - 2.0 * mp[2] - 6.0 * mp[5] + 6.0 * mp[6] - 9.0 * mp[7]) / 36.0;
        fp[5] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] + 6.0 * mp[3] + 3.0 * mp[4] + 6.0 * mp[5] + 3.0 * mp[6] + 9.0 * mp[8]) / 36.0;
        fp[6] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] - 6.0 * mp[3] - 3.0 * mp[4] + 6.0 * mp[5] + 3.0 * mp[6] - 9.0 * mp[8]) / 36.0;
        fp[7] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] - 6.0 * mp[3] - 3.0 * mp[4] - 6.0 * mp[5] - 3.0 * mp[6] + 9.0 * mp[8]) / 36.0;
        fp[8] = (4.0 * mp[0] + 2.0 * mp[1] + mp[2] + 6.0 * mp[3] + 3.0 * mp[4] - 6.0 * mp[5] - 3.0 * mp[6] - 9.0 * mp[8]) / 36.0;
    } else {
        fp[0] = (mp[0] - mp[1] + mp[2]) / 9.0;
        fp[1] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[3] / 6.0 - mp[4] / 6.0 + mp[7] * 0.25;
        fp[2] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[5] / 6.0 - mp[6] / 6.0 - mp[7] * 0.25;
        fp[3] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[3] / 6.0 + mp[4] / 6.0 + mp[7] * 0.25;
        fp[4] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[5] / 6.0 + mp[6] / 6.0 - mp[7] * 0.25;
        fp[5] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 + mp[8] * 0.25;
        fp[6] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 - mp[8] * 0.25;
        fp[7] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 + mp[8] * 0.25;
        fp[8] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 - mp[8] * 0.25;
    }
#else
    // throughput form: shared partial sums, the conserved moments (s = 0) pass through, constant reciprocals, FMA
    const double sa = (f[1] + f[3]) + (f[2] + f[4]), sd = (f[5] + f[7]) + (f[6] + f[8]);
    const double m0 = f[0] + (sa + sd);
    const double m1 = 2.0 * sd - sa - 4.0 * f[0], m2 = 4.0 * f[0] - 2.0 * sa + sd;
    const double ax = f[1] - f[3], dx = (f[5] - f[6]) + (f[8] - f[7]);
    const double ay = f[2] - f[4], dy = (f[5] + f[6]) - (f[7] + f[8]);
    const double m3 = ax + dx, m4 = dx - 2.0 * ax, m5 = ay + dy, m6 = dy - 2.0 * ay;
    const double m7 = (f[1] + f[3]) - (f[2] + f[4]), m8 = (f[5] + f[7]) - (f[6] + f[8]);
    const double uu = u * u, vv = v * v, q3 = 3.0 * (uu + vv);
    const double p1 = m1 - Snu * (m1 - (VARIANT == 2 ? q3 - 2.0 * rho : rho * (q3 - 2.0)));
    const double p2 = m2 - Snu * (m2 - (VARIANT == 2 ? rho - q3 : rho * (1.0 - q3)));
    const double p4 = m4 - Sq * (m4 + (VARIANT == 2 ? u : rho * u));
    const double p6 = m6 - Sq * (m6 + (VARIANT == 2 ? v : rho * v));
    const double p7 = m7 - Snu * (m7 - (VARIANT == 2 ? uu - vv : rho * (uu - vv)));
    const double p8 = m8 - Snu * (m8 - (VARIANT == 1 ? rho * (u * v) : u * v));
    constexpr double r9 = 1.0 / 9.0, r36 = 1.0 / 36.0, r6 = 1.0 / 6.0, r12 = 1.0 / 12.0;
    fp[0] = (m0 - p1 + p2) * r9;
    const double ca = (4.0 * m0 - p1 - 2.0 * p2) * r36, cd = (4.0 * m0 + 2.0 * p1 + p2) * r36;
    const double hx = (m3 - p4) * r6, hy = (m5 - p6) * r6, h7 = 0.25 * p7;
    fp[1] = ca + hx + h7; fp[3] = ca - hx + h7;
    fp[2] = ca + hy - h7; fp[4] = ca - hy - h7;
    const double gx = (2.0 * m3 + p4) * r12, gy = (2.0 * m5 + p6) * r12, h8 = 0.25 * p8;
    fp[5] = cd + gx + gy + h8; fp[6] = cd - gx + gy - h8;
    fp[7] = cd - gx - gy + h8; fp[8] = cd + gx - gy - h8;
#endif
}

// collision(): F (interior) + rho,u,v -> Fpost (interior)
template <int VARIANT>
__global__ void __launch_bounds__(128) k_l2_collision(Geom2 g, L2Params p, const double *__restrict__ F, const double *__restrict__ rho,
                                                      const double *__restrict__ u, const double *__restrict__ v,
                                                      double *__restrict__ Fpost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9], fp[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
    d2q9_collide<VARIANT>(f, rho[m], u[m], v[m], p.Snu, p.Sq, fp);
#pragma unroll
    for (int a = 0; a < 9; ++a) Fpost[a * g.sq + c] = fp[a];
}

// streaming() + bounceback()/boundary() + macro() of step n, collision() of step n+1.
// Unified boundary rule: a population whose upstream cell lies outside the GLOBAL box takes the opposite post-collision
// population of the cell itself (half-way bounce-back, bounceback.f90:7-29 == c:289-307); under the lid, populations 7 / 8
// also get - rho*(+U0)/6 / - rho*(-U0)/6 with rho of the previous macro() (bounceback.f90:34-36 == c:310-312), and because
// the top wall is processed last the lid term also holds in the two top corners.  Wall halos are never read.
// streaming() + bounceback() of one cell, then macro() (evolution.f90:105-107 == c:321-336; adds and IEEE divisions only:
// identical in both builds).  INC = L2I: the lid term is -(+-U0)/6 without rho (L2I:291-292) and u, v stay undivided (L2I:307-308).
template <bool INC>
__device__ __forceinline__ void l2_pull_macro(const Geom2 &g, const L2Params &p, const double *__restrict__ Fin,
                                              const double *__restrict__ rho_lid_in, int i, int j, double (&f)[9], double &rho,
                                              double &u, double &v) {
    const long long c = g.idx(0, i, j), sy = g.sy, sq = g.sq;
    const bool xp = g.wall[0] && i == g.nx, xm = g.wall[1] && i == 1, yp = g.wall[2] && j == g.ny, ym = g.wall[3] && j == 1;
#define L2_PULL(a, o, dx, dy)                                                                                       \
    {                                                                                                               \
        const bool wall_ = ((dx) == 1 && xm) || ((dx) == -1 && xp) || ((dy) == 1 && ym) || ((dy) == -1 && yp);     \
        f[a] = __ldg(Fin + (wall_ ? (o) * sq + c : (a) * sq + (c - (dy) * sy - (dx))));                             \
    }
    f[0] = __ldg(Fin + c);
    L2_PULL(1, 3, 1, 0) L2_PULL(2, 4, 0, 1) L2_PULL(3, 1, -1, 0) L2_PULL(4, 2, 0, -1)
    L2_PULL(5, 7, 1, 1) L2_PULL(6, 8, -1, 1) L2_PULL(7, 5, -1, -1) L2_PULL(8, 6, 1, -1)
#undef L2_PULL
    if (yp) {   // explicit _rn intrinsics: the fast build must not contract this
        if (INC) {
            f[7] = __dsub_rn(f[7], __ddiv_rn(p.U0, 6.0));
            f[8] = __dsub_rn(f[8], __ddiv_rn(-p.U0, 6.0));
        } else {
            const double r = rho_lid_in[i - 1];
            f[7] = __dsub_rn(f[7], __ddiv_rn(__dmul_rn(r, p.U0), 6.0));
            f[8] = __dsub_rn(f[8], __ddiv_rn(__dmul_rn(r, -p.U0), 6.0));
        }
    }
    rho = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(f[0], f[1]), f[2]), f[3]), f[4]), f[5]), f[6]), f[7]), f[8]);
    if (INC) {
        u = __dadd_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dsub_rn(f[1], f[3]), f[5]), f[6]), f[7]), f[8]);
        v = __dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(__dsub_rn(f[2], f[4]), f[5]), f[6]), f[7]), f[8]);
    } else {
        u = __ddiv_rn(__dadd_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dsub_rn(f[1], f[3]), f[5]), f[6]), f[7]), f[8]), rho);
        v = __ddiv_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(__dsub_rn(f[2], f[4]), f[5]), f[6]), f[7]), f[8]), rho);
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k_l2_fused(Geom2 g, L2Params p, const double *__restrict__ Fin, double *__restrict__ Fout,
                                                  const double *__restrict__ rho_lid_in, double *__restrict__ rho_lid_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    double f[9], fp[9], rho, u, v;
    l2_pull_macro<VARIANT == 2>(g, p, Fin, rho_lid_in, i, j, f, rho, u, v);
    if (g.wall[2] && j == g.ny) rho_lid_out[i - 1] = rho;
    d2q9_collide<VARIANT>(f, rho, u, v, p.Snu, p.Sq, fp);
    const long long c = g.idx(0, i, j);
#pragma unroll
    for (int a = 0; a < 9; ++a) Fout[a * g.sq + c] = fp[a];
}

// epilogue of a fused run: streaming() + bounceback() + macro() -> F (pre-collision) and the fields
template <bool INC>
__global__ void __launch_bounds__(128) k_l2_stream_macro(Geom2 g, L2Params p, const double *__restrict__ Fin, double *__restrict__ F,
                                                         const double *__restrict__ rho_lid_in, double *__restrict__ rho,
                                                         double *__restrict__ u, double *__restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    double f[9], r, uu, vv;
    l2_pull_macro<INC>(g, p, Fin, rho_lid_in, i, j, f, r, uu, vv);
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
#pragma unroll
    for (int a = 0; a < 9; ++a) F[a * g.sq + c] = f[a];
    rho[m] = r; u[m] = uu; v[m] = vv;
}

#ifndef MGLC_HOST_SHIM   // tests/host_shim/l2d_host.cpp runs the kernels above on the CPU and has no <<< >>>
int launch_l2_collision(const Geom2 &g, const L2Params &p, int variant, const double *F, const double *rho, const double *u,
                        const double *v, double *Fpost, cudaStream_t s) {
    const dim3 grid((g.nx + 127) / 128, g.ny);
    if (variant == 0) k_l2_collision<0><<<grid, 128, 0, s>>>(g, p, F, rho, u, v, Fpost);
    else if (variant == 1) k_l2_collision<1><<<grid, 128, 0, s>>>(g, p, F, rho, u, v, Fpost);
    else if (variant == 2) k_l2_collision<2><<<grid, 128, 0, s>>>(g, p, F, rho, u, v, Fpost);
    else k_l2_collision<3><<<grid, 128, 0, s>>>(g, p, F, rho, u, v, Fpost);
    return 1;
}
int launch_l2_stream_macro(const Geom2 &g, const L2Params &p, int variant, const double *Fin, double *F, const double *lid_in,
                           double *rho, double *u, double *v, cudaStream_t s) {
    const dim3 grid((g.nx + 127) / 128, g.ny);
    if (variant == 2) k_l2_stream_macro<true><<<grid, 128, 0, s>>>(g, p, Fin, F, lid_in, rho, u, v);
    else k_l2_stream_macro<false><<<grid, 128, 0, s>>>(g, p, Fin, F, lid_in, rho, u, v);
    return 1;
}
int launch_l2_fused(const Geom2 &g, const L2Params &p, int variant, const double *Fin, double *Fout, const double *lid_in,
                    double *lid_out, cudaStream_t s) {
    const dim3 grid((g.nx + 127) / 128, g.ny);
    if (variant == 0) k_l2_fused<0><<<grid, 128, 0, s>>>(g, p, Fin, Fout, lid_in, lid_out);
    else if (variant == 1) k_l2_fused<1><<<grid, 128, 0, s>>>(g, p, Fin, Fout, lid_in, lid_out);
    else if (variant == 2) k_l2_fused<2><<<grid, 128, 0, s>>>(g, p, Fin, Fout, lid_in, lid_out);
    else k_l2_fused<3><<<grid, 128, 0, s>>>(g, p, Fin, Fout, lid_in, lid_out);
    return 1;
}
#endif

}  // namespace MGLC_NS
}  // namespace mglc
