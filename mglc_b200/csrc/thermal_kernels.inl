// thermal_kernels.inl -- lattice-update kernels of the thermal double-distribution path (B3), included
// at the end of lbm_kernels.inl inside namespace mglc::MGLC_NS (so compiled twice: strict and fast).
//
// k_th_fused is the reference loop body (B3:222-248) rotated by half a step: streaming, bounceback,
// streamingT, bouncebackT, macro, macroT of step n and collision, collisionT of step n+1.  Per cell it
// reads 19 f + 7 g + Fx,Fy,Fz (the force the PREVIOUS collision computed, which this step's macro() needs,
// B3:998-1002) and writes the same 29 doubles: 464 B/cell (416 B of populations + 48 B of carried force).
// Walls are folded into the pull as address selection: f by half-way bounce-back on all six walls
// (B3:895-983), g by adiabatic bounce-back or the constant-temperature rule g_a = -g_opp + (6+paraA)/21 Tw
// (B3:1100-1210).  Both collisions use the fields left by the SAME macro()/macroT() (B3:226,234 both run
// before B3:242,244 of their own step), which is what the rotated kernel has in registers.

// the 7 temperature populations arriving at cell c
#define MGLC_PULL_G1(a, o, hit, face, off)                                                                 \
    {                                                                                                       \
        const double raw_ = __ldg(Gin + ((hit) ? (o) * sq + c : (a) * sq + c - (off)));                     \
        gq[a] = ((hit) && tp.bcT[face]) ? __dadd_rn(-raw_, tp.wallT[face]) : raw_;                          \
    }
#define MGLC_PULL_G()                                                                                       \
    gq[0] = __ldg(Gin + c);                                                                                 \
    MGLC_PULL_G1(1, 2, wf.xm, 1, 1)  MGLC_PULL_G1(2, 1, wf.xp, 0, -1)                                       \
    MGLC_PULL_G1(3, 4, wf.ym, 3, sy) MGLC_PULL_G1(4, 3, wf.yp, 2, -sy)                                      \
    MGLC_PULL_G1(5, 6, wf.zm, 5, sz) MGLC_PULL_G1(6, 5, wf.zp, 4, -sz)

__global__ void __launch_bounds__(128) k_th_collision(Geom g, ThermalParams tp, const double *__restrict__ F,
                                                      const double *__restrict__ rho, const double *__restrict__ u,
                                                      const double *__restrict__ v, const double *__restrict__ w,
                                                      const double *__restrict__ T, double *__restrict__ Fpost,
                                                      double *__restrict__ Fc) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    const long long n = (long long)g.nx * g.ny * g.nz;
    double f[19], fp[19];
#pragma unroll
    for (int a = 0; a < 19; ++a) f[a] = F[a * sq + c];
    const double r = rho[m], uu = u[m], vv = v[m], ww = w[m];
    double Fx, Fy, Fz;
    thermal_force(r, uu, vv, T[m], tp, Fx, Fy, Fz);
    Fc[m] = Fx; Fc[n + m] = Fy; Fc[2 * n + m] = Fz;
    d3q19_collide_thermal(f, r, uu, vv, ww, Fx, Fy, Fz, tp, fp);
#pragma unroll
    for (int a = 0; a < 19; ++a) Fpost[a * sq + c] = fp[a];
}

__global__ void __launch_bounds__(128) k_th_collisionT(Geom g, ThermalParams tp, const double *__restrict__ G,
                                                       const double *__restrict__ u, const double *__restrict__ v,
                                                       const double *__restrict__ w, const double *__restrict__ T,
                                                       double *__restrict__ Gpost) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    double gq[7], gp[7];
#pragma unroll
    for (int a = 0; a < 7; ++a) gq[a] = G[a * sq + c];
    d3q7_collide(gq, u[m], v[m], w[m], T[m], tp, gp);
#pragma unroll
    for (int a = 0; a < 7; ++a) Gpost[a * sq + c] = gp[a];
}

template <bool PEER>
__global__ void __launch_bounds__(128, 3) k_th_fused(Geom g, ThermalParams tp, const double *__restrict__ Fin,
                                                     double *__restrict__ Fout, const double *__restrict__ Gin,
                                                     double *__restrict__ Gout, const double *__restrict__ Fc_in,
                                                     double *__restrict__ Fc_out, int i0, int i1, int j0, int j1, int k0,
                                                     const PeerTable *__restrict__ pt) {
    // block (128,1) for full rows, (32,4) for the thin x-slabs of the boundary shell (see launch below)
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y, k = k0 + blockIdx.z;
    if (i > i1 || j > j1) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz;
    const long long c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    const long long n = (long long)g.nx * g.ny * g.nz;
    const WallFlags wf = wall_flags(g, i, j, k);
    double f[19], gq[7];
    MGLC_PULL_ALL();
    MGLC_PULL_G();
    double Fx = __ldg(Fc_in + m), Fy = __ldg(Fc_in + n + m), Fz = __ldg(Fc_in + 2 * n + m);
    double rho, u, v, w;
    d3q19_macro_forced(f, Fx, Fy, Fz, rho, u, v, w);          // macro() of step n with the force of step n's collision
    const double T = d3q7_temperature(gq);                    // macroT()
    {
        double gp[7];
        d3q7_collide(gq, u, v, w, T, tp, gp);                 // collisionT() of step n+1
#pragma unroll
        for (int a = 0; a < 7; ++a) Gout[a * sq + c] = gp[a];
        if (PEER && peer_cta_on_face(g, pt, i0 + (int)(blockIdx.x * blockDim.x), j0 + (int)(blockIdx.y * blockDim.y), k)) peer_store_g(pt, g, i, j, k, gp);
    }
    thermal_force(rho, u, v, T, tp, Fx, Fy, Fz);              // force of step n+1's collision
    Fc_out[m] = Fx; Fc_out[n + m] = Fy; Fc_out[2 * n + m] = Fz;
    double fp[19];
    d3q19_collide_thermal(f, rho, u, v, w, Fx, Fy, Fz, tp, fp);
#pragma unroll
    for (int a = 0; a < 19; ++a) Fout[a * sq + c] = fp[a];
    if (PEER) peer_store_f(pt, g, i, j, k, fp);
}

// leave the rotated state: pull f and g into the pre-collision lattices and write rho,u,v,w,T
__global__ void __launch_bounds__(128) k_th_stream_macro(Geom g, ThermalParams tp, const double *__restrict__ Fin,
                                                         double *__restrict__ F, const double *__restrict__ Gin,
                                                         double *__restrict__ G, const double *__restrict__ Fc_in,
                                                         double *__restrict__ rho_o, double *__restrict__ u_o,
                                                         double *__restrict__ v_o, double *__restrict__ w_o,
                                                         double *__restrict__ T_o) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz;
    const long long c = g.idx(0, i, j, k), m = g.cell(i, j, k);
    const long long n = (long long)g.nx * g.ny * g.nz;
    const WallFlags wf = wall_flags(g, i, j, k);
    double f[19], gq[7];
    MGLC_PULL_ALL();
    MGLC_PULL_G();
#pragma unroll
    for (int a = 0; a < 19; ++a) F[a * sq + c] = f[a];
#pragma unroll
    for (int a = 0; a < 7; ++a) G[a * sq + c] = gq[a];
    double rho, u, v, w;
    d3q19_macro_forced(f, Fc_in[m], Fc_in[n + m], Fc_in[2 * n + m], rho, u, v, w);
    rho_o[m] = rho; u_o[m] = u; v_o[m] = v; w_o[m] = w;
    T_o[m] = d3q7_temperature(gq);
}

#ifndef MGLC_HOST_SHIM
int launch_th_collision(const Geom &g, const ThermalParams &tp, const double *F, const double *rho, const double *u,
                        const double *v, const double *w, const double *T, double *Fpost, double *Fc, cudaStream_t s) {
    k_th_collision<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, tp, F, rho, u, v, w, T, Fpost, Fc);
    return 1;
}
int launch_th_collisionT(const Geom &g, const ThermalParams &tp, const double *G, const double *u, const double *v,
                         const double *w, const double *T, double *Gpost, cudaStream_t s) {
    k_th_collisionT<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, tp, G, u, v, w, T, Gpost);
    return 1;
}
int launch_th_fused(const Geom &g, const ThermalParams &tp, const double *Fin, double *Fout, const double *Gin,
                    double *Gout, const double *Fc_in, double *Fc_out, const int box[6], cudaStream_t s, const PeerTable *pt) {
    const int nxs = box[1] - box[0] + 1, nys = box[3] - box[2] + 1, nzs = box[5] - box[4] + 1;
    if (nxs <= 0 || nys <= 0 || nzs <= 0) return 0;
    const dim3 block = nxs <= 8 ? dim3(8, 16) : nxs <= 16 ? dim3(16, 8) : nxs <= 32 ? dim3(32, 4) : dim3(128, 1);
    const dim3 grid((nxs + block.x - 1) / block.x, (nys + block.y - 1) / block.y, nzs);
    if (pt) k_th_fused<true><<<grid, block, 0, s>>>(g, tp, Fin, Fout, Gin, Gout, Fc_in, Fc_out, box[0], box[1], box[2], box[3], box[4], pt);
    else k_th_fused<false><<<grid, block, 0, s>>>(g, tp, Fin, Fout, Gin, Gout, Fc_in, Fc_out, box[0], box[1], box[2], box[3], box[4], nullptr);
    return 1;
}
int launch_th_stream_macro(const Geom &g, const ThermalParams &tp, const double *Fin, double *F, const double *Gin,
                           double *G, const double *Fc_in, double *rho, double *u, double *v, double *w, double *T,
                           cudaStream_t s) {
    k_th_stream_macro<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, tp, Fin, F, Gin, G, Fc_in, rho, u, v, w, T);
    return 1;
}
#endif
