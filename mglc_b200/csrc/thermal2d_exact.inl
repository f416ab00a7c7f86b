// thermal2d_exact.inl -- the copy-type / order-preserving kernels of the 2-D thermal path (B2 = MPI/Buoyancy_driven_cavity/
// fortran/2d/mpi_blocked/): initial, streaming(T), bounceback(T), macro(T), check and calNuRe sums, halo pack/unpack, layout
// transposes.  Built once with -fmad=false inside thermal2d.cu's anonymous namespace; tests/host_shim/t2d_host.cpp includes the
// same text to run these kernels on the CPU against the oracle.

// module.F90:106-109
__constant__ int c_t2_ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
__constant__ int c_t2_ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
const int h_t2_ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
const int h_t2_ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
// populations leaving through each side, ascending = tag order of message_exchange.F90:10-59 (to right, left, top, bottom)
__constant__ int c_t2_face_pops[4][3] = {{1, 5, 8}, {3, 6, 7}, {2, 5, 6}, {4, 7, 8}};
// the one g population that crosses each side, message_exchange.F90:94-116
__constant__ int c_t2_face_popg[4] = {1, 3, 2, 4};

// initial(): initial.F90:199-212 (weights), :245-272 (fields), :276-288 (populations), :326-335
__global__ void __launch_bounds__(128) k_t2_initial(Geom2 g, T2Params p, int profile, int start, int total, double *__restrict__ F,
                                                    double *__restrict__ G, double *__restrict__ rho, double *__restrict__ u,
                                                    double *__restrict__ v, double *__restrict__ T, double *__restrict__ up,
                                                    double *__restrict__ vp, double *__restrict__ Tp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const double omega[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
    const double oT0 = (1.0 - p.paraA) / 5.0, oT1 = (p.paraA + 4.0) / 20.0;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    const double r = p.rho0;
    double uu = 0.0, vv = 0.0, t = 0.0;
    if (p.moving) {                                // the walls' velocities, RB2:466-481 (v on the vertical walls leaves the corners out)
        const int gi = p.start[0] + i, gj = p.start[1] + j;
        if (gj == p.total[1]) uu = t2_wall_u(p, true, gi);
        if (gj == 1) uu = t2_wall_u(p, false, gi);
        if (gj >= 2 && gj <= p.total[1] - 1) {
            if (gi == 1) vv = t2_wall_v(p, false, gj);
            if (gi == p.total[0]) vv = t2_wall_v(p, true, gj);
        }
    }
    // profile 1: VerticalWallsConstT (linear in x, :258); 2: HorizontalWallsConstT (linear in y, :268); start/total along that axis
    if (profile) t = (double)(start + (profile == 1 ? i : j) - 1) / (double)(total - 1) * (p.Tcold - p.Thot) + p.Thot;
    rho[m] = r; u[m] = uu; v[m] = vv; T[m] = t; up[m] = 0.0; vp[m] = 0.0; Tp[m] = 0.0;
    const double us2 = uu * uu + vv * vv;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        const double un = uu * (double)c_t2_ex[a] + vv * (double)c_t2_ey[a];
        F[a * g.sq + c] = r * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const double un = uu * (double)c_t2_ex[a] + vv * (double)c_t2_ey[a];
        G[a * g.sq + c] = t * (a == 0 ? oT0 : oT1) * (1.0 + 10.0 / (4.0 + p.paraA) * un);
    }
}

// streaming() evolution_f.F90:89-108 / streamingT() evolution_g.F90:49-68 (pull from the halo'd post-collision lattice; wall
// halos are read as they are, like the reference)
__global__ void __launch_bounds__(128) k_t2_streaming(Geom2 g, int nq, const double *__restrict__ Ppost, double *__restrict__ P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j);
    for (int a = 0; a < nq; ++a) P[a * g.sq + c] = Ppost[a * g.sq + c - c_t2_ey[a] * g.sy - c_t2_ex[a]];
}

// one thread per cell of the boundary ring: t -> (i, j)
__device__ __forceinline__ bool t2_ring_cell(const Geom2 &g, int t, int &i, int &j) {
    if (t < g.nx) { i = t + 1; j = 1; return true; }
    if (t < 2 * g.nx) { i = t - g.nx + 1; j = g.ny; return g.ny != 1; }
    if (t < 2 * g.nx + (g.ny - 2)) { i = 1; j = t - 2 * g.nx + 2; return true; }
    if (t < 2 * g.nx + 2 * (g.ny - 2)) { i = g.nx; j = t - 2 * g.nx - (g.ny - 2) + 2; return g.nx != 1; }
    return false;
}

// bounceback(): evolution_f.F90:283-321 (left, right, bottom, top; all half-way bounce-back)
// + moving walls (RB2:790-898): the diagonal populations off a wall get - rho*C/6, rho as the last macro() left it
__global__ void __launch_bounds__(128) k_t2_bounceback(Geom2 g, T2Params p, const double *__restrict__ Fpost, double *__restrict__ F,
                                                       const double *__restrict__ rho) {
    int i, j;
    if (!t2_ring_cell(g, blockIdx.x * blockDim.x + threadIdx.x, i, j)) return;
    const long long c = g.idx(0, i, j), sq = g.sq;
    const int perx = p.perx;
    if (perx) {   // VerticalWallsPeriodicalU, seq/bouyancy2d_acc.F90:777-791 (before the horizontal walls, like the reference)
        const long long wrap = g.nx - 1;
        if (i == 1) { F[1 * sq + c] = Fpost[1 * sq + c + wrap]; F[5 * sq + c] = Fpost[5 * sq + c + wrap]; F[8 * sq + c] = Fpost[8 * sq + c + wrap]; }
        if (i == g.nx) { F[3 * sq + c] = Fpost[3 * sq + c - wrap]; F[6 * sq + c] = Fpost[6 * sq + c - wrap]; F[7 * sq + c] = Fpost[7 * sq + c - wrap]; }
    }
    if (g.wall[1] && i == 1) { F[1 * sq + c] = Fpost[3 * sq + c]; F[5 * sq + c] = Fpost[7 * sq + c]; F[8 * sq + c] = Fpost[6 * sq + c]; }
    if (g.wall[0] && i == g.nx) { F[3 * sq + c] = Fpost[1 * sq + c]; F[6 * sq + c] = Fpost[8 * sq + c]; F[7 * sq + c] = Fpost[5 * sq + c]; }
    if (g.wall[3] && j == 1) { F[2 * sq + c] = Fpost[4 * sq + c]; F[5 * sq + c] = Fpost[7 * sq + c]; F[6 * sq + c] = Fpost[8 * sq + c]; }
    if (g.wall[2] && j == g.ny) { F[4 * sq + c] = Fpost[2 * sq + c]; F[7 * sq + c] = Fpost[5 * sq + c]; F[8 * sq + c] = Fpost[6 * sq + c]; }
    if (!p.moving) return;
    const bool xp = g.wall[0] && i == g.nx, xm = g.wall[1] && i == 1, yp = g.wall[2] && j == g.ny, ym = g.wall[3] && j == 1;
    const int gi = p.start[0] + i, gj = p.start[1] + j;
    const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    for (int a = 5; a < 9; ++a) {
        const int ex = c_t2_ex[a], ey = c_t2_ey[a];
        const bool hx = (ex == 1 && xm) || (ex == -1 && xp), hy = (ey == 1 && ym) || (ey == -1 && yp);
        if (hx | hy) F[a * sq + c] = Fpost[opp[a] * sq + c] - rho[g.cell(i, j)] * t2_wall_coef(p, ex, ey, hx, hy, gi, gj) / 6.0;
    }
}

// bouncebackT(): evolution_g.F90:79-142 (adiabatic: g(a) = g_post(opp); constant T: g(a) = -g_post(opp) + (4+paraA)/10*T_wall)
__global__ void __launch_bounds__(128) k_t2_bouncebackT(Geom2 g, T2Params p, const double *__restrict__ Gpost, double *__restrict__ G) {
    int i, j;
    if (!t2_ring_cell(g, blockIdx.x * blockDim.x + threadIdx.x, i, j)) return;
    const long long c = g.idx(0, i, j), sq = g.sq;
    if (p.perx) {   // VerticalWallsPeriodicalT, seq/bouyancy2d_acc.F90:1037-1045
        if (i == 1) G[1 * sq + c] = Gpost[1 * sq + c + (g.nx - 1)];
        if (i == g.nx) G[3 * sq + c] = Gpost[3 * sq + c - (g.nx - 1)];
    }
    if (g.wall[3] && j == 1) G[2 * sq + c] = p.bcT[3] ? -Gpost[4 * sq + c] + p.wallT[3] : Gpost[4 * sq + c];
    if (g.wall[2] && j == g.ny) G[4 * sq + c] = p.bcT[2] ? -Gpost[2 * sq + c] + p.wallT[2] : Gpost[2 * sq + c];
    if (g.wall[1] && i == 1) G[1 * sq + c] = p.bcT[1] ? -Gpost[3 * sq + c] + p.wallT[1] : Gpost[3 * sq + c];
    if (g.wall[0] && i == g.nx) G[3 * sq + c] = p.bcT[0] ? -Gpost[1 * sq + c] + p.wallT[0] : Gpost[1 * sq + c];
    // RB2:1086-1106: in a corner cell the population off the vertical wall takes the constant-temperature rule of the plate
    const bool ym = g.wall[3] && j == 1, yp = g.wall[2] && j == g.ny;
    if (p.cornersT && (ym | yp)) {
        const int plate = ym ? 3 : 2;
        if (p.bcT[plate]) {
            if (g.wall[1] && i == 1) G[1 * sq + c] = -Gpost[3 * sq + c] + p.wallT[plate];
            if (g.wall[0] && i == g.nx) G[3 * sq + c] = -Gpost[1 * sq + c] + p.wallT[plate];
        }
    }
}

// macro(): evolution_f.F90:328-342
__global__ void __launch_bounds__(128) k_t2_macro(Geom2 g, const double *__restrict__ F, const double *__restrict__ Fx,
                                                  const double *__restrict__ Fy, double *__restrict__ rho, double *__restrict__ u,
                                                  double *__restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
    const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    rho[m] = r;
    u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8] + 0.5 * Fx[m]) / r;
    v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8] + 0.5 * Fy[m]) / r;
}
// macroT(): evolution_g.F90:163-176
__global__ void __launch_bounds__(128) k_t2_macroT(Geom2 g, const double *__restrict__ G, double *__restrict__ T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j);
    T[g.cell(i, j)] = G[c] + G[g.sq + c] + G[2 * g.sq + c] + G[3 * g.sq + c] + G[4 * g.sq + c];
}

#ifndef MGLC_HOST_SHIM   // block reductions use shared memory and __syncthreads(): not part of the sequential host emulation
// check(): check.F90:10-39: error1 = sum (du^2 + dv^2), error2 = sum (u^2 + v^2), error5 = sum |dT|, error6 = sum |T|;
// up, vp, Tp <- u, v, T.   Fixed grid and tree: reproducible summation order.
constexpr int T2_RED_BLOCKS = 296;
constexpr int T2_RED_WORDS = 4;
__device__ __forceinline__ void t2_block_reduce(double (&e)[T2_RED_WORDS], int nwords, double *__restrict__ part) {
    __shared__ double sh[T2_RED_WORDS][256];
    for (int q = 0; q < nwords; ++q) sh[q][threadIdx.x] = e[q];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int q = 0; q < nwords; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int q = 0; q < nwords; ++q) part[T2_RED_WORDS + T2_RED_WORDS * blockIdx.x + q] = sh[q][0];
}
__global__ void __launch_bounds__(256) k_t2_check_partial(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                          const double *__restrict__ T, double *__restrict__ up, double *__restrict__ vp,
                                                          double *__restrict__ Tp, double *__restrict__ part) {
    double e[T2_RED_WORDS] = {0.0, 0.0, 0.0, 0.0};
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const double a = u[q], b = v[q], t = T[q];
        const double da = a - up[q], db = b - vp[q];
        e[0] += da * da + db * db;
        e[1] += a * a + b * b;
        e[2] += fabs(t - Tp[q]);
        e[3] += fabs(t);
        up[q] = a; vp[q] = b; Tp[q] = t;
    }
    t2_block_reduce(e, 4, part);
}
// calNuRe()'s volume sums, NuRe.F90:27-78 with global indices: sum (i-nxHalf)*v - (j-nyHalf)*u, sum v*T, sum u*u + v*v
__global__ void __launch_bounds__(256) k_t2_nure_partial(int nx, int ny, int sx, int sy, int nxHalf, int nyHalf,
                                                         const double *__restrict__ u, const double *__restrict__ v,
                                                         const double *__restrict__ T, double *__restrict__ part) {
    double e[T2_RED_WORDS] = {0.0, 0.0, 0.0, 0.0};
    const long long n = (long long)nx * ny;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nx) + 1, j = (int)(q / nx) + 1;
        const double a = u[q], b = v[q];
        e[0] += (double)(sx + i - nxHalf) * b - (double)(sy + j - nyHalf) * a;
        e[1] += b * T[q];
        e[2] += a * a + b * b;
    }
    t2_block_reduce(e, 3, part);
}
__global__ void k_t2_reduce_final(int nblocks, int nwords, double *__restrict__ part) {
    for (int q = 0; q < nwords; ++q) {
        double e = 0.0;
        for (int b = 0; b < nblocks; ++b) e += part[T2_RED_WORDS + T2_RED_WORDS * b + q];
        part[q] = e;
    }
}

#endif

// halo messages.  f: message_passing_f(), message_exchange.F90:1-79 -- dir 0..3 = to right(+x), left(-x), top(+y), bottom(-y),
// three populations over the interior range, buffer [slot][t]; dir 4..7 = the corner population 5..8 crosses, one value.
// g: message_passing_g(), :85-118 -- dir 8..11 = the same four sides, one population, no corners.
__device__ __forceinline__ void t2_msg_cell(const Geom2 &g, int dir, int ghost, int t, int &i, int &j) {
    if (dir >= 8) dir -= 8;
    if (dir < 4) {
        const int axis = dir >> 1, plus = !(dir & 1);
        const int nfix = axis == 0 ? g.nx : g.ny;
        const int fix = ghost ? (plus ? 0 : nfix + 1) : (plus ? nfix : 1);
        i = axis == 0 ? fix : 1 + t;
        j = axis == 1 ? fix : 1 + t;
    } else {
        const int a = dir + 1, px = c_t2_ex[a] > 0, py = c_t2_ey[a] > 0;
        i = ghost ? (px ? 0 : g.nx + 1) : (px ? g.nx : 1);
        j = ghost ? (py ? 0 : g.ny + 1) : (py ? g.ny : 1);
    }
}
__device__ __forceinline__ int t2_msg_pop(int dir, int slot) {
    return dir >= 8 ? c_t2_face_popg[dir - 8] : dir < 4 ? c_t2_face_pops[dir][slot] : dir + 1;
}
__global__ void k_t2_pack(Geom2 g, const double *__restrict__ Ppost, int dir, int n1, int npop, double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * npop) return;
    int i, j;
    t2_msg_cell(g, dir, 0, t % n1, i, j);
    buf[t] = Ppost[g.idx(t2_msg_pop(dir, t / n1), i, j)];
}
__global__ void k_t2_unpack(Geom2 g, double *__restrict__ Ppost, int dir, int n1, int npop, const double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * npop) return;
    int i, j;
    t2_msg_cell(g, dir, 1, t % n1, i, j);
    Ppost[g.idx(t2_msg_pop(dir, t / n1), i, j)] = buf[t];
}

// reference layout (population index fastest; with_halo: (0:q-1,0:nx+1,0:ny+1), else (0:q-1,nx,ny)) <-> SoA rows
__global__ void __launch_bounds__(128) k_t2_aos_to_soa(Geom2 g, int nq, const double *__restrict__ aos, double *__restrict__ P, int with_halo) {
    const int w = with_halo ? g.nx + 2 : g.nx, hgt = with_halo ? g.ny + 2 : g.ny;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= hgt) return;
    const int i = with_halo ? x : x + 1, j = with_halo ? y : y + 1;
    const long long src = (long long)nq * (x + (long long)w * y), c = g.idx(0, i, j);
    for (int a = 0; a < nq; ++a) P[a * g.sq + c] = aos[src + a];
}
__global__ void __launch_bounds__(128) k_t2_soa_to_aos(Geom2 g, int nq, const double *__restrict__ P, double *__restrict__ aos, int with_halo) {
    const int w = with_halo ? g.nx + 2 : g.nx, hgt = with_halo ? g.ny + 2 : g.ny;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= hgt) return;
    const int i = with_halo ? x : x + 1, j = with_halo ? y : y + 1;
    const long long dst = (long long)nq * (x + (long long)w * y), c = g.idx(0, i, j);
    for (int a = 0; a < nq; ++a) aos[dst + a] = P[a * g.sq + c];
}

