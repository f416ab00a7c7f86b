// thermal2d_kernels.inl -- 2-D thermal D2Q9 + D2Q5 kernels; compiled twice (thermal2d.cu: namespace strict, -fmad=false;
// thermal2d_fast.cu: namespace fast, -fmad=true).   B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/
// Device layout: SoA F[a][j][x] (9 populations) and G[a][j][x] (5), one-cell halo ring, rows padded like the 3-D lattice
// (interior cell i = 1 at x-index OX, pitch a multiple of 16 doubles): thread <-> cell, threadIdx.x along x, every warp store
// 128-byte aligned.  The fused kernel is the reference loop body (main.F90:84-108) rotated by half a step: streaming() +
// bounceback() + streamingT() + bouncebackT() + macro() + macroT() of step n and collision() + collisionT() of step n+1 in one
// pass: (9 + 5) loads + (9 + 5) stores of fp64 + the carried force Fy (8 B in, 8 B out) = 240 B per cell.
#include "thermal2d.cuh"

namespace mglc {
namespace MGLC_NS {

#include "d2q9_thermal.inl"

// collision(): F (interior) + rho,u,v,T -> Fpost (interior), Fx = 0, Fy          evolution_f.F90:1-84
__global__ void __launch_bounds__(128) k_t2_collision(Geom2 g, T2Params p, const double *__restrict__ F, const double *__restrict__ rho,
                                                      const double *__restrict__ u, const double *__restrict__ v,
                                                      const double *__restrict__ T, double *__restrict__ Fpost, double *__restrict__ Fx,
                                                      double *__restrict__ Fy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9], fp[9], fy;
#pragma unroll
    for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
    if (p.variant) t2_collide<true>(f, rho[m], u[m], v[m], T[m], p.Snu, p.Sq, p.gBeta, p.Tref, fp, fy);
    else t2_collide<false>(f, rho[m], u[m], v[m], T[m], p.Snu, p.Sq, p.gBeta, p.Tref, fp, fy);
#pragma unroll
    for (int a = 0; a < 9; ++a) Fpost[a * g.sq + c] = fp[a];
    Fx[m] = 0.0; Fy[m] = fy;
}

// collisionT(): G (interior) + u,v,T -> Gpost (interior)                          evolution_g.F90:1-46
__global__ void __launch_bounds__(128) k_t2_collisionT(Geom2 g, T2Params p, const double *__restrict__ G, const double *__restrict__ u,
                                                       const double *__restrict__ v, const double *__restrict__ T,
                                                       double *__restrict__ Gpost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double gg[5], gp[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) gg[a] = G[a * g.sq + c];
    t2_collideT(gg, u[m], v[m], T[m], p.Qd, p.Qnu, p.paraA, gp);
#pragma unroll
    for (int a = 0; a < 5; ++a) Gpost[a * g.sq + c] = gp[a];
}

// streaming() + bounceback() and streamingT() + bouncebackT() of one cell, then macro() + macroT().
// Unified boundary rule: a population whose upstream cell lies outside the GLOBAL box takes the opposite post-collision
// population of the cell itself -- f: half-way bounce-back on all four no-slip walls (evolution_f.F90:283-321; every wall
// assigns the same value, so their order does not matter); g: the same for an adiabatic wall, and -g_post(opp) +
// (4+paraA)/10*T_wall for a constant-temperature wall (evolution_g.F90:79-142).  The wall test only selects the load
// address; wall halos are never read.
// Moving walls (p.moving, the sheared Rayleigh-Benard programs): the diagonal populations off a wall get - rho_prev*C/6 with
// rho of the previous macro() (rho_prev: the caller reads it from the field array, whose wall cells every launch keeps current).
__device__ __forceinline__ void t2_pull_macro(const Geom2 &g, const T2Params &p, const double *__restrict__ Fin,
                                              const double *__restrict__ Gin, double Fy, int i, int j, double rho_prev, double (&f)[9],
                                              double (&gg)[5], double &rho, double &u, double &v, double &T) {
    const long long c = g.idx(0, i, j), sy = g.sy, sq = g.sq;
    const bool xp = g.wall[0] && i == g.nx, xm = g.wall[1] && i == 1, yp = g.wall[2] && j == g.ny, ym = g.wall[3] && j == 1;
    // periodic vertical walls (acc:777-791): a population entering through x takes the SAME population from the SAME row of
    // the opposite column (also the diagonal ones); the horizontal walls are processed afterwards and win in the corner cells
    const bool pxm = p.perx && i == 1, pxp = p.perx && i == g.nx;
    const long long wrap = g.nx - 1;
#define T2_PULL(a, o, dx, dy)                                                                                       \
    {                                                                                                               \
        const bool wall_ = ((dx) == 1 && xm) || ((dx) == -1 && xp) || ((dy) == 1 && ym) || ((dy) == -1 && yp);     \
        const bool per_ = ((dx) == 1 && pxm) || ((dx) == -1 && pxp);                                                \
        f[a] = __ldg(Fin + (wall_ ? (o) * sq + c : per_ ? (a) * sq + (c + (dx) * wrap) : (a) * sq + (c - (dy) * sy - (dx)))); \
    }
    f[0] = __ldg(Fin + c);
    T2_PULL(1, 3, 1, 0) T2_PULL(2, 4, 0, 1) T2_PULL(3, 1, -1, 0) T2_PULL(4, 2, 0, -1)
    T2_PULL(5, 7, 1, 1) T2_PULL(6, 8, -1, 1) T2_PULL(7, 5, -1, -1) T2_PULL(8, 6, 1, -1)
#undef T2_PULL
    if (p.moving && (xp | xm | yp | ym)) {
        const int gi = p.start[0] + i, gj = p.start[1] + j;
#define T2_MOVE(a, dx, dy)                                                                                          \
    {                                                                                                               \
        const bool hx_ = ((dx) == 1 && xm) || ((dx) == -1 && xp), hy_ = ((dy) == 1 && ym) || ((dy) == -1 && yp);   \
        if (hx_ | hy_) f[a] = __dsub_rn(f[a], __ddiv_rn(__dmul_rn(rho_prev, t2_wall_coef(p, dx, dy, hx_, hy_, gi, gj)), 6.0)); \
    }
        T2_MOVE(5, 1, 1) T2_MOVE(6, -1, 1) T2_MOVE(7, -1, -1) T2_MOVE(8, 1, -1)
#undef T2_MOVE
    }
    // side: 0 = +x wall (population 3 comes off it), 1 = -x (1), 2 = +y (4), 3 = -y (2)
#define T2_PULLG(a, o, off, hit, side, per, pwrap)                                                                  \
    {                                                                                                               \
        const double raw_ = __ldg(Gin + ((hit) ? (o) * sq + c : (per) ? (a) * sq + (c + (pwrap)) : (a) * sq + (c - (off)))); \
        gg[a] = ((hit) && p.bcT[side]) ? __dadd_rn(-raw_, p.wallT[side]) : raw_;                                    \
    }
    gg[0] = __ldg(Gin + c);
    // cornersT (RB2:1086-1106): in a corner cell the population off the vertical wall takes the rule of the plate it touches
    const int plate = ym ? 3 : 2;
    const bool cornerT = p.cornersT && (ym | yp) && p.bcT[plate];
    const int sxm = (cornerT && xm) ? plate : 1, sxp = (cornerT && xp) ? plate : 0;
    T2_PULLG(1, 3, 1, xm, sxm, pxm, wrap) T2_PULLG(2, 4, sy, ym, 3, false, 0) T2_PULLG(3, 1, -1, xp, sxp, pxp, -wrap)
    T2_PULLG(4, 2, -sy, yp, 2, false, 0)
#undef T2_PULLG
    t2_macro_cell(f, gg, 0.0, Fy, rho, u, v, T);
}

__global__ void __launch_bounds__(128) k_t2_fused(Geom2 g, T2Params p, const double *__restrict__ Fin, double *__restrict__ Fout,
                                                  const double *__restrict__ Gin, double *__restrict__ Gout, double *__restrict__ Fy,
                                                  double *__restrict__ rho_field) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9], gg[5], fp[9], gp[5], rho, u, v, T, fy;
    const bool wall_cell = p.moving && ((g.wall[0] && i == g.nx) | (g.wall[1] && i == 1) | (g.wall[2] && j == g.ny) | (g.wall[3] && j == 1));
    t2_pull_macro(g, p, Fin, Gin, Fy[m], i, j, wall_cell ? rho_field[m] : 0.0, f, gg, rho, u, v, T);
    if (wall_cell) rho_field[m] = rho;           // what the next bounceback() of this cell multiplies the wall velocity with
    if (p.variant) t2_collide<true>(f, rho, u, v, T, p.Snu, p.Sq, p.gBeta, p.Tref, fp, fy);
    else t2_collide<false>(f, rho, u, v, T, p.Snu, p.Sq, p.gBeta, p.Tref, fp, fy);
    t2_collideT(gg, u, v, T, p.Qd, p.Qnu, p.paraA, gp);
#pragma unroll
    for (int a = 0; a < 9; ++a) Fout[a * g.sq + c] = fp[a];
#pragma unroll
    for (int a = 0; a < 5; ++a) Gout[a * g.sq + c] = gp[a];
    Fy[m] = fy;
}

// epilogue of a fused run: the pulls + macro() + macroT() -> F, G (pre-collision) and the fields
__global__ void __launch_bounds__(128) k_t2_stream_macro(Geom2 g, T2Params p, const double *__restrict__ Fin, double *__restrict__ F,
                                                         const double *__restrict__ Gin, double *__restrict__ G,
                                                         const double *__restrict__ Fy, double *__restrict__ rho, double *__restrict__ u,
                                                         double *__restrict__ v, double *__restrict__ T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9], gg[5], r, uu, vv, tt;
    t2_pull_macro(g, p, Fin, Gin, Fy[m], i, j, p.moving ? rho[m] : 0.0, f, gg, r, uu, vv, tt);
#pragma unroll
    for (int a = 0; a < 9; ++a) F[a * g.sq + c] = f[a];
#pragma unroll
    for (int a = 0; a < 5; ++a) G[a * g.sq + c] = gg[a];
    rho[m] = r; u[m] = uu; v[m] = vv; T[m] = tt;
}

#ifndef MGLC_HOST_SHIM   // tests/host_shim/t2d_host.cpp runs the kernels above on the CPU and has no <<< >>>
static inline dim3 t2_grid(const Geom2 &g) { return dim3((g.nx + 127) / 128, g.ny); }

int launch_t2_collision(const Geom2 &g, const T2Params &p, const double *F, const double *rho, const double *u, const double *v,
                        const double *T, double *Fpost, double *Fx, double *Fy, cudaStream_t s) {
    k_t2_collision<<<t2_grid(g), 128, 0, s>>>(g, p, F, rho, u, v, T, Fpost, Fx, Fy);
    return 1;
}
int launch_t2_collisionT(const Geom2 &g, const T2Params &p, const double *G, const double *u, const double *v, const double *T,
                         double *Gpost, cudaStream_t s) {
    k_t2_collisionT<<<t2_grid(g), 128, 0, s>>>(g, p, G, u, v, T, Gpost);
    return 1;
}
int launch_t2_fused(const Geom2 &g, const T2Params &p, const double *Fin, double *Fout, const double *Gin, double *Gout, double *Fy,
                    double *rho, cudaStream_t s) {
    k_t2_fused<<<t2_grid(g), 128, 0, s>>>(g, p, Fin, Fout, Gin, Gout, Fy, rho);
    return 1;
}
int launch_t2_stream_macro(const Geom2 &g, const T2Params &p, const double *Fin, double *F, const double *Gin, double *G,
                           const double *Fy, double *rho, double *u, double *v, double *T, cudaStream_t s) {
    k_t2_stream_macro<<<t2_grid(g), 128, 0, s>>>(g, p, Fin, F, Gin, G, Fy, rho, u, v, T);
    return 1;
}
#endif

}  // namespace MGLC_NS
}  // namespace mglc
