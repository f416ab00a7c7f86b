// lid2d_exact.inl -- the copy-type / order-preserving kernels of the 2-D lid-driven cavity path (L2C, L2F, L2I of lid2d.cu):
// initial, streaming, bounceback, macro, the lid row, check sums, halo pack/unpack, layout transposes.  Built once with
// -fmad=false inside lid2d.cu's anonymous namespace; tests/host_shim/l2d_host.cpp includes the same text to run these kernels
// on the CPU against the oracle.

// commondata.f90:25-27 == c:24-25
__constant__ int c_ex9[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
__constant__ int c_ey9[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
const int h_ex9[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
const int h_ey9[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
// populations leaving through each side, ascending = tag order of ex_sendrecv.f90:9-45 (to right, left, top, bottom)
__constant__ int c_face_pops9[4][3] = {{1, 5, 8}, {3, 6, 7}, {2, 5, 6}, {4, 7, 8}};

// initial(): initial.f90:40-66 == c:123-151; INC = L2I:137-163 (rho = 0, f = omega*(...))
template <bool INC>
__global__ void __launch_bounds__(128) k_l2_initial(Geom2 g, L2Params p, int lid, double *__restrict__ F, double *__restrict__ rho,
                                                    double *__restrict__ u, double *__restrict__ v, double *__restrict__ up,
                                                    double *__restrict__ vp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const double omega[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    const double r = p.rho0, uu = (lid && j == g.ny) ? p.U0 : 0.0, vv = 0.0;
    rho[m] = INC ? 0.0 : r; u[m] = uu; v[m] = vv; up[m] = 0.0; vp[m] = 0.0;
    const double us2 = uu * uu + vv * vv;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        const double un = uu * (double)c_ex9[a] + vv * (double)c_ey9[a];
        if (INC) F[a * g.sq + c] = omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
        else F[a * g.sq + c] = r * omega[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
    }
}

// streaming(): evolution.f90:80-97 (pull from the halo'd f_post; wall halos are read as they are, like the reference)
__global__ void __launch_bounds__(128) k_l2_streaming(Geom2 g, const double *__restrict__ Fpost, double *__restrict__ F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j);
#pragma unroll
    for (int a = 0; a < 9; ++a) F[a * g.sq + c] = Fpost[a * g.sq + c - c_ey9[a] * g.sy - c_ex9[a]];
}

// bounceback(): bounceback.f90:7-40 == boundary(), c:286-313.  One thread per wall cell applies left, right, bottom, top in
// the reference's order, so the later wall wins in the corners exactly as in the sequential loops.  INC = L2I:266-293.
template <bool INC>
__global__ void __launch_bounds__(128) k_l2_bounceback(Geom2 g, L2Params p, const double *__restrict__ Fpost,
                                                       const double *__restrict__ rho, double *__restrict__ F) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (t < g.nx) { i = t + 1; j = 1; }
    else if (t < 2 * g.nx) { i = t - g.nx + 1; j = g.ny; if (g.ny == 1) return; }
    else if (t < 2 * g.nx + (g.ny - 2)) { i = 1; j = t - 2 * g.nx + 2; }
    else if (t < 2 * g.nx + 2 * (g.ny - 2)) { i = g.nx; j = t - 2 * g.nx - (g.ny - 2) + 2; if (g.nx == 1) return; }
    else return;
    const long long c = g.idx(0, i, j), sq = g.sq;
    if (g.wall[1] && i == 1) { F[1 * sq + c] = Fpost[3 * sq + c]; F[5 * sq + c] = Fpost[7 * sq + c]; F[8 * sq + c] = Fpost[6 * sq + c]; }
    if (g.wall[0] && i == g.nx) { F[3 * sq + c] = Fpost[1 * sq + c]; F[6 * sq + c] = Fpost[8 * sq + c]; F[7 * sq + c] = Fpost[5 * sq + c]; }
    if (g.wall[3] && j == 1) { F[2 * sq + c] = Fpost[4 * sq + c]; F[5 * sq + c] = Fpost[7 * sq + c]; F[6 * sq + c] = Fpost[8 * sq + c]; }
    if (g.wall[2] && j == g.ny) {
        F[4 * sq + c] = Fpost[2 * sq + c];
        if (INC) {
            F[7 * sq + c] = Fpost[5 * sq + c] - p.U0 / 6.0;
            F[8 * sq + c] = Fpost[6 * sq + c] - (-p.U0) / 6.0;
        } else {
            const double r = rho[g.cell(i, j)];
            F[7 * sq + c] = Fpost[5 * sq + c] - r * p.U0 / 6.0;
            F[8 * sq + c] = Fpost[6 * sq + c] - r * (-p.U0) / 6.0;
        }
    }
}

// macro(): evolution.f90:105-113 == c:321-336; INC = L2I:304-310 (u, v undivided)
template <bool INC>
__global__ void __launch_bounds__(128) k_l2_macro(Geom2 g, const double *__restrict__ F, double *__restrict__ rho,
                                                  double *__restrict__ u, double *__restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.nx) return;
    const long long c = g.idx(0, i, j), m = g.cell(i, j);
    double f[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) f[a] = F[a * g.sq + c];
    const double r = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    rho[m] = r;
    if (INC) {
        u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]);
        v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]);
    } else {
        u[m] = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / r;
        v[m] = (f[2] - f[4] + f[5] + f[6] - f[7] - f[8]) / r;
    }
}

// the lid row of rho, kept beside the rotated loop (the lid term uses rho of the previous macro(), bounceback.f90:35-36)
__global__ void k_l2_lid_row(Geom2 g, const double *__restrict__ rho, double *__restrict__ lid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i <= g.nx) lid[i - 1] = rho[g.cell(i, g.ny)];
}

#ifndef MGLC_HOST_SHIM   // shared-memory tree reduction: not emulated by the sequential host sweep
// check(): evolution.f90:128-147 == c:341-363: error1 = sum (du^2 + dv^2), error2 = sum (u^2 + v^2); up, vp <- u, v
// INC = L2I:325-332: error1 = sum sqrt(du^2 + dv^2), error2 = sum sqrt(u^2 + v^2)
constexpr int L2_CHECK_BLOCKS = 296;   // fixed: reproducible summation order
template <bool INC>
__global__ void __launch_bounds__(256) k_l2_check_partial(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                          double *__restrict__ up, double *__restrict__ vp, double *__restrict__ part) {
    double e1 = 0.0, e2 = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const double a = u[q], b = v[q];
        const double da = a - up[q], db = b - vp[q];
        if (INC) { e1 += sqrt(da * da + db * db); e2 += sqrt(a * a + b * b); }
        else { e1 += da * da + db * db; e2 += a * a + b * b; }
        up[q] = a; vp[q] = b;
    }
    __shared__ double s1[256], s2[256];
    s1[threadIdx.x] = e1; s2[threadIdx.x] = e2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 + 2 * blockIdx.x] = s1[0]; part[3 + 2 * blockIdx.x] = s2[0]; }
}
__global__ void k_l2_check_final(int nblocks, double *__restrict__ part) {
    double e1 = 0.0, e2 = 0.0;
    for (int b = 0; b < nblocks; ++b) { e1 += part[2 + 2 * b]; e2 += part[3 + 2 * b]; }
    part[0] = e1; part[1] = e2;
}
#endif

// halo messages of message_passing_sendrecv(), ex_sendrecv.f90:9-78: dir 0..3 = to right(+x), left(-x), top(+y), bottom(-y), three
// populations over the interior range, buffer [slot][t]; dir 4..7 = the corner population 5..8 crosses, one value.
__device__ __forceinline__ void l2_msg_cell(const Geom2 &g, int dir, int ghost, int t, int &i, int &j) {
    if (dir < 4) {
        const int axis = dir >> 1, plus = !(dir & 1);
        const int nfix = axis == 0 ? g.nx : g.ny;
        const int fix = ghost ? (plus ? 0 : nfix + 1) : (plus ? nfix : 1);
        i = axis == 0 ? fix : 1 + t;
        j = axis == 1 ? fix : 1 + t;
    } else {
        const int a = dir + 1, px = c_ex9[a] > 0, py = c_ey9[a] > 0;
        i = ghost ? (px ? 0 : g.nx + 1) : (px ? g.nx : 1);
        j = ghost ? (py ? 0 : g.ny + 1) : (py ? g.ny : 1);
    }
}
__global__ void k_l2_pack(Geom2 g, const double *__restrict__ Fpost, int dir, int n1, int npop, double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * npop) return;
    int i, j;
    l2_msg_cell(g, dir, 0, t % n1, i, j);
    const int a = dir < 4 ? c_face_pops9[dir][t / n1] : dir + 1;
    buf[t] = Fpost[g.idx(a, i, j)];
}
__global__ void k_l2_unpack(Geom2 g, double *__restrict__ Fpost, int dir, int n1, int npop, const double *__restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 * npop) return;
    int i, j;
    l2_msg_cell(g, dir, 1, t % n1, i, j);
    const int a = dir < 4 ? c_face_pops9[dir][t / n1] : dir + 1;
    Fpost[g.idx(a, i, j)] = buf[t];
}

// reference layout (population index fastest; with_halo: (0:8,0:nx+1,0:ny+1), else (0:8,nx,ny)) <-> SoA rows
__global__ void __launch_bounds__(128) k_l2_aos_to_soa(Geom2 g, const double *__restrict__ aos, double *__restrict__ F, int with_halo) {
    const int w = with_halo ? g.nx + 2 : g.nx, hgt = with_halo ? g.ny + 2 : g.ny;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= hgt) return;
    const int i = with_halo ? x : x + 1, j = with_halo ? y : y + 1;
    const long long src = 9LL * (x + (long long)w * y), c = g.idx(0, i, j);
#pragma unroll
    for (int a = 0; a < 9; ++a) F[a * g.sq + c] = aos[src + a];
}
__global__ void __launch_bounds__(128) k_l2_soa_to_aos(Geom2 g, const double *__restrict__ F, double *__restrict__ aos, int with_halo) {
    const int w = with_halo ? g.nx + 2 : g.nx, hgt = with_halo ? g.ny + 2 : g.ny;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= hgt) return;
    const int i = with_halo ? x : x + 1, j = with_halo ? y : y + 1;
    const long long dst = 9LL * (x + (long long)w * y), c = g.idx(0, i, j);
#pragma unroll
    for (int a = 0; a < 9; ++a) aos[dst + a] = F[a * g.sq + c];
}
