// decomp.cpp -- host-only part of the C ABI: the reference's 3-D Cartesian decomposition and halo plan
// (MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/main.f90:24-72,144-212 and ex_sendrecv.f90),
// computed without MPI.  Nothing here touches a GPU, so it is testable on a CPU-only machine.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>

#include "../../include/mglc.h"

namespace mglc {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
static const int ex[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int ey[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int ez[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int face_pops[6][5] = {{1, 7, 9, 11, 13}, {2, 8, 10, 12, 14}, {3, 7, 8, 15, 17},
                                    {4, 9, 10, 16, 18}, {5, 11, 12, 15, 16}, {6, 13, 14, 17, 18}};
}  // namespace mglc
using mglc::set_error;

extern "C" {

int mglc_version(void) { return MGLC_VERSION; }
const char *mglc_last_error(void) { return mglc::g_err; }
const char *mglc_strerror(int code) {
    switch (code) {
        case MGLC_OK: return "ok";
        case MGLC_E_INVALID: return "invalid argument";
        case MGLC_E_CUDA: return "CUDA error";
        case MGLC_E_NCCL: return "NCCL error";
        case MGLC_E_NOMEM: return "out of memory";
        case MGLC_E_STATE: return "invalid state for this call";
        case MGLC_E_NOGPU: return "no CUDA device (libmglc has no CPU path)";
        case MGLC_E_DIVERGED: return "device-side fatal physics flag";
        default: return "unknown error";
    }
}

int mglc_dims_create(int nranks, int dims[3]) {
    if (nranks < 1 || !dims) { set_error("mglc_dims_create: nranks=%d", nranks); return MGLC_E_INVALID; }
    // balanced factorisation a >= b >= c, smallest a first, then smallest b (what MPI_Dims_create
    // returns for the rank counts the reference runs with: 2,4,6,8,12,24 -- P4/mpi.sh:8-19)
    int best[3] = {nranks, 1, 1};
    for (int a = 1; a <= nranks; ++a) {
        if (nranks % a) continue;
        for (int b = 1; b <= a; ++b) {
            if ((nranks / a) % b) continue;
            const int c = nranks / a / b;
            if (c > b) continue;
            if (a < best[0] || (a == best[0] && b < best[1])) { best[0] = a; best[1] = b; best[2] = c; }
        }
    }
    memcpy(dims, best, sizeof best);
    return MGLC_OK;
}

int mglc_decompose_1d(int total_n, int rank, int nranks, int *local_n, int *start) {
    if (total_n < 1 || nranks < 1 || rank < 0 || rank >= nranks || !local_n) {
        set_error("mglc_decompose_1d: total_n=%d rank=%d nranks=%d", total_n, rank, nranks);
        return MGLC_E_INVALID;
    }
    const int n = total_n / nranks, m = total_n % nranks;
    *local_n = n + (rank < m ? 1 : 0);
    if (start) *start = rank * n + (rank < m ? rank : m);
    return MGLC_OK;
}

int mglc_cart_rank(const int dims[3], const int coords[3], int *rank) {
    if (!dims || !coords || !rank) return MGLC_E_INVALID;
    for (int d = 0; d < 3; ++d)
        if (coords[d] < 0 || coords[d] >= dims[d]) { *rank = -1; return MGLC_OK; }   // MPI_PROC_NULL
    *rank = (coords[0] * dims[1] + coords[1]) * dims[2] + coords[2];
    return MGLC_OK;
}

int mglc_cart_coords(const int dims[3], int rank, int coords[3]) {
    if (!dims || !coords || rank < 0 || rank >= dims[0] * dims[1] * dims[2]) {
        set_error("mglc_cart_coords: rank=%d", rank);
        return MGLC_E_INVALID;
    }
    coords[2] = rank % dims[2];
    coords[1] = (rank / dims[2]) % dims[1];
    coords[0] = rank / (dims[2] * dims[1]);
    return MGLC_OK;
}

int mglc_cart_neighbors(const int dims[3], const int coords[3], int nbr_surface[6], int nbr_line[12]) {
    if (!dims || !coords || !nbr_surface || !nbr_line) return MGLC_E_INVALID;
    for (int d = 0; d < 3; ++d) {
        int p[3] = {coords[0], coords[1], coords[2]}, m[3] = {coords[0], coords[1], coords[2]};
        p[d] += 1; m[d] -= 1;
        mglc_cart_rank(dims, p, &nbr_surface[2 * d]);
        mglc_cart_rank(dims, m, &nbr_surface[2 * d + 1]);
    }
    for (int a = 7; a < 19; ++a) {
        const int c[3] = {coords[0] + mglc::ex[a], coords[1] + mglc::ey[a], coords[2] + mglc::ez[a]};
        mglc_cart_rank(dims, c, &nbr_line[a - 7]);
    }
    return MGLC_OK;
}

int mglc_relaxation_rates(double tau, double *Snu, double *Sq) {
    if (!(tau > 0.5) || !Snu || !Sq) { set_error("mglc_relaxation_rates: tau=%g must exceed 0.5", tau); return MGLC_E_INVALID; }
    *Snu = 1.0 / tau;                                        // L3/commondata.f90:42
    *Sq = 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0);
    return MGLC_OK;
}

int mglc_lbm_desc_init(mglc_lbm_desc *d, const int gn[3], const int dims_or_zero[3], int nranks, int rank,
                       double reynolds, double U0, double rho0) {
    if (!d || !gn || nranks < 1 || rank < 0 || rank >= nranks) { set_error("mglc_lbm_desc_init: bad arguments"); return MGLC_E_INVALID; }
    memset(d, 0, sizeof *d);
    d->lattice = MGLC_D3Q19; d->collision = MGLC_MRT_LID; d->arith = MGLC_ARITH_FAST; d->kernel = MGLC_KERNEL_AUTO;
    memcpy(d->gn, gn, 3 * sizeof(int));
    if (dims_or_zero && dims_or_zero[0] > 0) {
        memcpy(d->dims, dims_or_zero, 3 * sizeof(int));
        if (d->dims[0] * d->dims[1] * d->dims[2] != nranks) { set_error("mglc_lbm_desc_init: dims do not multiply to nranks"); return MGLC_E_INVALID; }
    } else {
        mglc_dims_create(nranks, d->dims);
    }
    mglc_cart_coords(d->dims, rank, d->coords);
    for (int q = 0; q < 3; ++q) {
        if (gn[q] < d->dims[q]) { set_error("mglc_lbm_desc_init: fewer cells than ranks along dim %d", q); return MGLC_E_INVALID; }
        mglc_decompose_1d(gn[q], d->coords[q], d->dims[q], &d->ln[q], &d->start[q]);
    }
    d->tau = U0 * (double)gn[0] / reynolds * 3.0 + 0.5;      // L3/commondata.f90:9
    d->U0 = U0; d->rho0 = rho0; d->device = 0;
    return MGLC_OK;
}

int mglc_thermal_desc_init(mglc_lbm_desc *d, const int gn[3], const int dims_or_zero[3], int nranks, int rank,
                           double rayleigh, double prandtl, double mach, double ekman) {
    int rc = mglc_lbm_desc_init(d, gn, dims_or_zero, nranks, rank, 1.0, 0.0, 1.0);
    if (rc) return rc;
    if (!(rayleigh > 0.0) || !(prandtl > 0.0) || !(ekman > 0.0)) { set_error("mglc_thermal_desc_init: Ra, Pr, Ek must be positive"); return MGLC_E_INVALID; }
    d->lattice = MGLC_D3Q19_D3Q7; d->collision = MGLC_MRT_THERMAL;
    const double nz = (double)gn[2];
    // module commondata, B3:33-47,73-74 (same operation order)
    d->tau = 0.5 + mach * nz * sqrt(3.0 * prandtl / rayleigh);
    const double viscosity = (d->tau - 0.5) / 3.0;
    const double diffusivity = viscosity / prandtl;
    d->omegaRot = viscosity / 2.0 / ekman / (double)(gn[2] * gn[2]);
    d->paraA = 42.0 * sqrt(3.0) * diffusivity - 6.0;
    const double gBeta1 = rayleigh * viscosity * diffusivity / nz;
    d->gBeta = gBeta1 / (double)(gn[2] * gn[2]);
    d->Thot = 1.0; d->Tcold = 0.0; d->Tref = 0.0;
    d->Qd = 3.0 - sqrt(3.0);
    d->Qnu = 4.0 * sqrt(3.0) - 6.0;
    d->U0 = 0.0; d->rho0 = 1.0;
    // benchmarkCavity: LeftRightWallsConstT (hot at j = 1), TopBottomPlatesAdiabatic, BackFrontWallsAdiabatic (B3:15-19)
    const int shipped[6] = {MGLC_BCT_ADIABATIC, MGLC_BCT_ADIABATIC, MGLC_BCT_CONST_COLD, MGLC_BCT_CONST_HOT,
                            MGLC_BCT_ADIABATIC, MGLC_BCT_ADIABATIC};
    memcpy(d->bcT, shipped, sizeof shipped);
    return MGLC_OK;
}

int mglc_halo_plan(const mglc_lbm_desc *d, mglc_halo_msg msgs[18], int *nmsgs) {
    if (!d || !msgs || !nmsgs) return MGLC_E_INVALID;
    int ns[6], nl[12];
    mglc_cart_neighbors(d->dims, d->coords, ns, nl);
    const int n[3] = {d->ln[0], d->ln[1], d->ln[2]};
    int k = 0;
    for (int dir = 0; dir < 6; ++dir) {                       // ex_sendrecv.f90:12-59
        const int axis = dir >> 1;
        const int area = (axis == 0) ? n[1] * n[2] : (axis == 1 ? n[0] * n[2] : n[0] * n[1]);
        mglc_halo_msg &m = msgs[k++];
        m.dir = dir; m.npop = 5;
        m.send_to = ns[dir];
        m.recv_from = ns[dir ^ 1];                            // data travelling in +x arrives from my -x neighbour
        m.send_count = m.send_to >= 0 ? 5 * area : 0;
        m.recv_count = m.recv_from >= 0 ? 5 * area : 0;       // tangential sizes agree across a Cartesian face
        memcpy(m.pops, mglc::face_pops[dir], sizeof m.pops);
    }
    static const int order[12] = {7, 10, 9, 8, 11, 14, 13, 12, 15, 18, 17, 16};   // ex_sendrecv.f90:64-123
    static const int opp[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    for (int q = 0; q < 12; ++q) {
        const int a = order[q];
        const int len = (mglc::ex[a] == 0) ? n[0] : (mglc::ey[a] == 0 ? n[1] : n[2]);
        mglc_halo_msg &m = msgs[k++];
        m.dir = a; m.npop = 1;
        m.send_to = nl[a - 7];
        m.recv_from = nl[opp[a] - 7];
        m.send_count = m.send_to >= 0 ? len : 0;
        m.recv_count = m.recv_from >= 0 ? len : 0;
        m.pops[0] = a; m.pops[1] = m.pops[2] = m.pops[3] = m.pops[4] = -1;
    }
    *nmsgs = k;
    return MGLC_OK;
}

// The 2-D drivers' messages (Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/ex_sendrecv.f90:9-78 and
// Buoyancy_driven_cavity/fortran/2d/mpi_blocked/message_exchange.F90:1-118) for the rank at `rank` of a dims[0] x dims[1] grid
// (rank = c0*dims[1] + c1): msgs 0..3 = f faces to right(+x), left(-x), top(+y), bottom(-y), three populations over the interior
// range; 4..7 = the corner population 5..8 crosses, one value; 8..11 = the thermal driver's g faces, one population each.
int mglc_halo_plan_2d(int total_nx, int total_ny, const int dims[2], int rank, mglc_halo_msg msgs[12], int *nmsgs) {
    if (!dims || !msgs || !nmsgs || dims[0] < 1 || dims[1] < 1 || rank < 0 || rank >= dims[0] * dims[1]) return MGLC_E_INVALID;
    static const int ex9[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, ey9[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    static const int face_pops[4][3] = {{1, 5, 8}, {3, 6, 7}, {2, 5, 6}, {4, 7, 8}}, face_popg[4] = {1, 3, 2, 4};
    const int c0 = rank / dims[1], c1 = rank % dims[1];
    int n[2], start;
    mglc_decompose_1d(total_nx, c0, dims[0], &n[0], &start);
    mglc_decompose_1d(total_ny, c1, dims[1], &n[1], &start);
    auto cart = [&](int a, int b) { return (a < 0 || a >= dims[0] || b < 0 || b >= dims[1]) ? -1 : a * dims[1] + b; };
    for (int dir = 0; dir < 12; ++dir) {
        mglc_halo_msg &m = msgs[dir];
        const int d = dir >= 8 ? dir - 8 : dir;
        const int ox = d < 4 ? (d == 0) - (d == 1) : ex9[d + 1], oy = d < 4 ? (d == 2) - (d == 3) : ey9[d + 1];
        const int n1 = d < 4 ? ((d >> 1) == 0 ? n[1] : n[0]) : 1;
        m.dir = dir;
        m.npop = d < 4 ? (dir >= 8 ? 1 : 3) : 1;
        m.send_to = cart(c0 + ox, c1 + oy);
        m.recv_from = cart(c0 - ox, c1 - oy);          // what travels in +x arrives from my -x neighbour
        m.send_count = m.send_to >= 0 ? n1 * m.npop : 0;
        m.recv_count = m.recv_from >= 0 ? n1 * m.npop : 0;      // the extent along a face is shared by both sides of a Cartesian cut
        for (int q = 0; q < 5; ++q) m.pops[q] = -1;
        if (dir >= 8) m.pops[0] = face_popg[d];
        else if (d < 4) for (int q = 0; q < 3; ++q) m.pops[q] = face_pops[d][q];
        else m.pops[0] = d + 1;
    }
    *nmsgs = 12;
    return MGLC_OK;
}

}  // extern "C"
