// lbm_aa_exact.inl -- the order-preserving kernels of the AA-pattern path (lbm_aa.cu), built once with -fmad=false; the same
// text runs on the CPU in tests/host_shim/aa_host.cpp.  Needs lbm_aa_kernels.inl's AA_PULL_ALL / AA_LID_PULL and d3q19_macro.

// POST layout -> rho,u,v,w: streaming() + bounceback() + macro() of the current loop body without touching the lattice
// (the epilogue of a run that ended on an even launch, and what mglc_aa_download_macro needs in that state)
__global__ void __launch_bounds__(128) k_aa_macro_post(Geom g, LbmParams p, const double *A, const double *rho_lid_in, double *rho_o,
                                                       double *__restrict__ u_o, double *__restrict__ v_o, double *__restrict__ w_o) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > g.nx) return;
    const long long sq = g.sq, sy = g.sy, sz = g.sz, c = g.idx(0, i, j, k);
    const AaWalls wf = aa_walls(g, i, j, k);
    const double *__restrict__ Ain = A;
    double f[19];
    AA_PULL_ALL();
    AA_LID_PULL(rho_lid_in);
    double rho, u, v, w;
    d3q19_macro(f, rho, u, v, w);
    const long long m = g.cell(i, j, k);
    rho_o[m] = rho; u_o[m] = u; v_o[m] = v; w_o[m] = w;
}

// POST layout -> f(0:18, cells c0 .. c0+ncells) in the reference's layout (population index fastest): the pre-collision
// populations of the current loop body, gathered chunk by chunk for mglc_aa_download_f
__global__ void __launch_bounds__(128) k_aa_gather_f(Geom g, LbmParams p, const double *A, const double *rho_lid_in, long long c0,
                                                     long long ncells, double *__restrict__ aos) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= ncells) return;
    const long long cell = c0 + q;
    const int i = (int)(cell % g.nx) + 1, j = (int)((cell / g.nx) % g.ny) + 1, k = (int)(cell / ((long long)g.nx * g.ny)) + 1;
    const long long sq = g.sq, sy = g.sy, sz = g.sz, c = g.idx(0, i, j, k);
    const AaWalls wf = aa_walls(g, i, j, k);
    const double *__restrict__ Ain = A;
    double f[19];
    AA_PULL_ALL();
    AA_LID_PULL(rho_lid_in);
#pragma unroll
    for (int a = 0; a < 19; ++a) aos[19 * q + a] = f[a];
}

// the lid plane (k = nz) of rho, kept beside the lattice: the moving-lid term uses rho of the previous macro()
__global__ void k_aa_lid_plane(Geom g, const double *__restrict__ rho, double *__restrict__ lid) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x, n = (long long)g.nx * g.ny;
    if (q < n) lid[q] = rho[n * (g.nz - 1) + q];
}
