/*
 * mglc.h -- C ABI of libmglc.so: the B200-native replacement for the per-timestep hot path of
 * cheryli/MGLC (collide + stream lattice update, boundary kernels, 3-D subdomain halo exchange).
 *
 * The reference has no FFI today: its drivers call argument-less Fortran subroutines over
 * `module commondata` globals (MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/main.f90:85-103,
 * "L3" below).  This header defines the seam a driver binds through ISO_C_BINDING
 * (fortran/mglc_iso_c.f90).  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *  - every function returns MGLC_OK (0) or a negative MGLC_E_* code; nothing aborts or throws;
 *    mglc_last_error() returns a thread-local message for the last failure.
 *  - all array pointers are HOST pointers to the reference's Fortran layout (column-major, the
 *    population index fastest): f(0:18,nx,ny,nz), f_post(0:18,0:nx+1,0:ny+1,0:nz+1),
 *    rho,u,v,w(nx,ny,nz) (L3/initial.f90:35-44).  The library copies; it keeps no caller pointer.
 *  - one handle = one subdomain on one GPU, used by one host thread at a time (the reference's own
 *    rule: one MPI rank, strictly sequential).
 *  - there is NO CPU fallback: without a CUDA device the device entry points return MGLC_E_NOGPU.
 */
#ifndef MGLC_H
#define MGLC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGLC_VERSION 105

/* ---- status codes ---- */
#define MGLC_OK          0
#define MGLC_E_INVALID  (-1)   /* bad argument / descriptor */
#define MGLC_E_CUDA     (-2)   /* CUDA runtime error */
#define MGLC_E_NCCL     (-3)   /* NCCL error */
#define MGLC_E_NOMEM    (-4)   /* device or host allocation failed */
#define MGLC_E_STATE    (-5)   /* call not valid in the handle's current state */
#define MGLC_E_NOGPU    (-6)   /* no CUDA device: the product has no CPU path */
#define MGLC_E_DIVERGED (-7)   /* device-side fatal-physics flag (reference: stop / MPI_Abort) */

/* ---- enums (plain ints in the ABI) ---- */
enum { MGLC_D3Q19 = 0, MGLC_D3Q19_D3Q7 = 1, MGLC_D2Q9 = 2 };
/* collision operator: MGLC_MRT_LID reproduces L3/collision.f90 including the meq(12) quirk (:85) */
enum { MGLC_MRT_LID = 0, MGLC_MRT_THERMAL = 1, MGLC_BGK = 2 };
/* arithmetic: STRICT keeps the reference's expression order, true divisions and no FMA contraction
 * (bit-identical to the CPU oracle); FAST uses reciprocal multiplies + FMA (<= 1e-12 rel. L2). */
enum { MGLC_ARITH_FAST = 0, MGLC_ARITH_STRICT = 1 };
/* how halos move between subdomains */
enum { MGLC_TRANSPORT_NONE = 0, MGLC_TRANSPORT_LOCAL = 1, MGLC_TRANSPORT_NCCL = 2 };
/* thermal wall kinds of bouncebackT() (B3 = MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90:1100-1210):
 * adiabatic g_a = g_post_opp, constant temperature g_a = -g_post_opp + (6+paraA)/21 * Thot|Tcold */
enum { MGLC_BCT_ADIABATIC = 0, MGLC_BCT_CONST_HOT = 1, MGLC_BCT_CONST_COLD = 2,
       MGLC_BCT_PERIODIC = 3 /* 2-D thermal driver only: both vertical walls, seq/bouyancy2d_acc.F90:15,22 */ };
/* fused-kernel variant (MGLC_KERNEL_AUTO picks the fastest validated one) */
enum { MGLC_KERNEL_AUTO = 0, MGLC_KERNEL_DIRECT = 1, MGLC_KERNEL_TMA = 2 };

typedef struct mglc_lbm   mglc_lbm;     /* opaque: one subdomain on one GPU */
typedef struct mglc_comm  mglc_comm;    /* opaque: NCCL communicator wrapper */
typedef struct mglc_group mglc_group;   /* opaque: P subdomains driven by one process */

/* Plain-old-data descriptor; replaces the compile-time `parameter`s of `module commondata`
 * (L3/commondata.f90:4-15,42-53). */
typedef struct mglc_lbm_desc {
    int lattice;        /* MGLC_D3Q19 ...                                              */
    int collision;      /* MGLC_MRT_LID ...                                            */
    int arith;          /* MGLC_ARITH_FAST | MGLC_ARITH_STRICT                         */
    int kernel;         /* MGLC_KERNEL_*                                               */
    int gn[3];          /* total_nx, total_ny, total_nz          (commondata.f90:4)    */
    int dims[3];        /* Cartesian process grid, dims(0) <-> x (main.f90:24)         */
    int coords[3];      /* my coordinates in it                  (main.f90:36)         */
    int ln[3];          /* local nx, ny, nz                      (main.f90:37-39)      */
    int start[3];       /* 0-based global offset of local cell 1                       */
    double tau;         /* tauf                                  (commondata.f90:9)    */
    double U0;          /* lid speed                             (commondata.f90:8)    */
    double rho0;        /* initial density                       (commondata.f90:7)    */
    int device;         /* CUDA device ordinal                                         */
    /* ---- MGLC_D3Q19_D3Q7 only (module commondata of B3:26-47,73-74; BC macros B3:5-19) ---- */
    int bcT[6];         /* thermal wall kind per face +x,-x,+y,-y,+z,-z : MGLC_BCT_*       */
    int reserved[1];
    double paraA;       /* 42*sqrt(3)*diffusivity - 6                                  */
    double gBeta;       /* Ra*nu*kappa/total_nz^3                                      */
    double Tref, Thot, Tcold;
    double omegaRot;    /* omegaRatating = nu/2/Ekman/total_nz^2 (Coriolis)            */
    double Qd, Qnu;     /* D3Q7 relaxation rates 3-sqrt(3), 4*sqrt(3)-6                */
} mglc_lbm_desc;

/* one halo message of message_passing_sendrecv() (L3/ex_sendrecv.f90): dir 0..5 = faces
 * +x,-x,+y,-y,+z,-z (5 populations each), dir 7..18 = the edge that population `dir` crosses. */
typedef struct mglc_halo_msg {
    int dir;            /* direction id as above                                       */
    int send_to;        /* rank receiving my outgoing data, -1 = MPI_PROC_NULL         */
    int recv_from;      /* rank whose data lands in my opposite halo, -1 = none        */
    int npop;           /* 5 for faces, 1 for edges                                    */
    int send_count;     /* doubles I send  (npop * my slab size)                       */
    int recv_count;     /* doubles I receive (npop * sender's slab size)               */
    int pops[5];        /* population indices, ascending (tag order of ex_sendrecv.f90) */
} mglc_halo_msg;

/* ================= host-only helpers (work without a GPU) ================= */
int         mglc_version(void);
const char *mglc_last_error(void);
const char *mglc_strerror(int code);
int         mglc_device_count(int *n);

/* MPI_Dims_create(np,3,dims) with dims=0 -- L3/main.f90:24 */
int mglc_dims_create(int nranks, int dims[3]);
/* decompose_1d -- L3/main.f90:144-155 (first total_n mod nranks ranks get one extra plane) */
int mglc_decompose_1d(int total_n, int rank, int nranks, int *local_n, int *start);
/* MPI_Cart_rank / MPI_Cart_coords of the row-major communicator -- L3/main.f90:25,36 */
int mglc_cart_rank(const int dims[3], const int coords[3], int *rank /* -1 outside */);
int mglc_cart_coords(const int dims[3], int rank, int coords[3]);
/* MPI_Cart_shift x3 + MPI_Cart_find_corners -- L3/main.f90:43-47,158-212.
 * nbr_surface[0..5] = +x,-x,+y,-y,+z,-z ; nbr_line[0..11] = populations 7..18 ; -1 = PROC_NULL */
int mglc_cart_neighbors(const int dims[3], const int coords[3], int nbr_surface[6], int nbr_line[12]);
/* fill a descriptor the way L3/main.f90:24-39 + commondata.f90:9 do: dims (if dims[0]==0) from
 * nranks, coords/ln/start from rank, tau = U0*total_nx/Re*3+0.5 */
int mglc_lbm_desc_init(mglc_lbm_desc *d, const int gn[3], const int dims_or_zero[3], int nranks,
                       int rank, double reynolds, double U0, double rho0);
/* thermal descriptor the way B3 does it: decomposition as above; tau = 0.5 + Ma*total_nz*sqrt(3 Pr/Ra),
 * paraA, gBeta, omegaRot, Qd, Qnu from B3:33-47,73-74; Thot = 1, Tcold = Tref = 0; the shipped
 * benchmarkCavity wall set (y walls constant T, hot at j = 1; x and z walls adiabatic; all walls no-slip) */
int mglc_thermal_desc_init(mglc_lbm_desc *d, const int gn[3], const int dims_or_zero[3], int nranks, int rank,
                           double rayleigh, double prandtl, double mach, double ekman);
/* the 18 messages of one message_passing_sendrecv() for this subdomain, in the reference's order */
int mglc_halo_plan(const mglc_lbm_desc *d, mglc_halo_msg msgs[18], int *nmsgs);
/* the 2-D drivers' messages for the rank at `rank` (= c0*dims[1] + c1): 0..3 f faces (+x,-x,+y,-y; 3 populations), 4..7 the corner
 * population 5..8 crosses (1 value), 8..11 the thermal driver's g faces (1 population) -- 2d_revised/mpi_blocked/ex_sendrecv.f90:9-78,
 * Buoyancy_driven_cavity/fortran/2d/mpi_blocked/message_exchange.F90:1-118; buffer layout [slot][t] over the sender's interior range */
int mglc_halo_plan_2d(int total_nx, int total_ny, const int dims[2], int rank, mglc_halo_msg msgs[12], int *nmsgs);
/* the message tables the 2-D CUDA drivers build for themselves at create time (host-only views, for the CPU suite: they must
 * equal the plan above; `pops` is not filled) */
int mglc_l2d_msg_table(int total_nx, int total_ny, const int dims[2], int rank, mglc_halo_msg out[8]);
int mglc_t2d_msg_table(int total_nx, int total_ny, const int dims[2], int rank, mglc_halo_msg out[12]);
/* relaxation rates Snu, Sq -- L3/commondata.f90:42 */
int mglc_relaxation_rates(double tau, double *Snu, double *Sq);

/* ================= communicator (one process per GPU) ================= */
/* ncclGetUniqueId / ncclCommInitRank; the id is broadcast by the caller (MPI_Bcast in a Fortran
 * driver, torch.distributed in the Python harness).  Replaces MPI_Cart_create, L3/main.f90:25. */
int mglc_comm_unique_id(char id[128]);
int mglc_comm_init_rank(mglc_comm **c, const char id[128], int nranks, int rank, int device);
int mglc_comm_destroy(mglc_comm *c);

/* ================= one subdomain ================= */
/* allocate(f, f_post, rho,u,v,w,up,vp,wp) -- L3/initial.f90:35-44 ; comm may be NULL (1 rank) */
int mglc_lbm_create(mglc_lbm **h, const mglc_lbm_desc *d, mglc_comm *comm_or_null);
int mglc_lbm_destroy(mglc_lbm *h);                       /* deallocate -- L3/main.f90:123-131 */
/* initial(): rho0, u=U0 on the lid plane, f = feq -- L3/initial.f90:46-73 (device-side) */
int mglc_lbm_initial(mglc_lbm *h);
int mglc_lbm_upload(mglc_lbm *h, const double *f, const double *rho, const double *u,
                    const double *v, const double *w);   /* any pointer may be NULL = keep */
int mglc_lbm_upload_fpost(mglc_lbm *h, const double *f_post);    /* with halo, for bit-exact tests */
int mglc_lbm_download_macro(mglc_lbm *h, double *rho, double *u, double *v, double *w); /* output(), L3/output.f90:12-119 */
int mglc_lbm_download_f(mglc_lbm *h, double *f);         /* backupData()-style checkpoints */
int mglc_lbm_download_fpost(mglc_lbm *h, double *f_post);

/* one call per reference subroutine (a driver can swap ONE call at a time) */
int mglc_collision(mglc_lbm *h);    /* collision()                 L3/collision.f90:1-205   */
int mglc_exchange(mglc_lbm *h);     /* message_passing_sendrecv()  L3/ex_sendrecv.f90:1-129 */
int mglc_streaming(mglc_lbm *h);    /* streaming()                 L3/streaming.f90:1-23    */
int mglc_bounceback(mglc_lbm *h);   /* bounceback()                L3/bounce_back.f90:1-86  */
int mglc_macro(mglc_lbm *h);        /* macro()                     L3/macro.f90:1-28        */
int mglc_check(mglc_lbm *h, double *errorU);   /* check() incl. Allreduce  L3/check.f90:1-37 */

/* thermal double-distribution handles (lattice == MGLC_D3Q19_D3Q7): mglc_collision / mglc_exchange /
 * mglc_streaming / mglc_bounceback / mglc_macro above act on f as B3's collision (+force, B3:640-864),
 * f_message_passing_sendrecv, streaming, bounceback (no-slip on all walls), macro (+F/2, B3:986-1010);
 * the temperature populations have their own five subroutines */
int mglc_collisionT(mglc_lbm *h);   /* collisionT()                  B3:1014-1070 */
int mglc_exchange_g(mglc_lbm *h);   /* g_message_passing_sendrecv()  B3:1421-1468 */
int mglc_streamingT(mglc_lbm *h);   /* streamingT()                  B3:1075-1098 */
int mglc_bouncebackT(mglc_lbm *h);  /* bouncebackT()                 B3:1100-1210 */
int mglc_macroT(mglc_lbm *h);       /* macroT()                      B3:1215-1232 */
int mglc_check_thermal(mglc_lbm *h, double *errorU, double *errorT);   /* check()  B3:1236-1283 */
/* g (0:6,nx,ny,nz), g_post (0:6,0:nx+1,..), T, Fx,Fy,Fz (nx,ny,nz); any pointer may be NULL = skip */
int mglc_lbm_upload_thermal(mglc_lbm *h, const double *g, const double *T, const double *Fx, const double *Fy, const double *Fz);
int mglc_lbm_download_thermal(mglc_lbm *h, double *g, double *T, double *Fx, double *Fy, double *Fz);
int mglc_lbm_upload_gpost(mglc_lbm *h, const double *g_post);
int mglc_lbm_download_gpost(mglc_lbm *h, double *g_post);

/* in-loop diagnostics on the device: no full-field download needed during long runs.
 * calNuRe(): NuVolAvg = sum(w T)/N * total_nz/diffusivity + 1, ReVolAvg = sqrt(sum(u^2+v^2+w^2)/N) * total_nz/viscosity over the
 * global box, viscosity = (tau-1/2)/3, diffusivity = viscosity/Prandtl -- B3/mpi_blocked/RaNu.F90:13-47, module.F90:38-39 */
int mglc_calNuRe(mglc_lbm *h, double prandtl, double *NuVolAvg, double *ReVolAvg);
/* one line of rho|u|v|w|T (field 0..4) along `axis` through global 1-based (g1, g2) on the two other axes, e.g. getVelocity()'s
 * u(nxHalf,nyHalf,:) and w(:,nyHalf,nzHalf), L3/output.f90:334-344.  out holds ln[axis] values starting at global index *first;
 * *count = 0 when the line does not cross this subdomain */
int mglc_lbm_download_line(mglc_lbm *h, int field, int axis, int g1, int g2, double *out, int *first, int *count);

/* fused fast path == nsteps iterations of the loop body L3/main.f90:85-97 (B3:222-248 for thermal handles)
 * (collision, exchange, streaming, bounceback, macro); prologue/epilogue handled inside, the state
 * afterwards (f, f_post, rho,u,v,w) is the reference's state after the same number of iterations */
int mglc_lbm_step(mglc_lbm *h, int nsteps);
int mglc_lbm_sync(mglc_lbm *h);
/* same, bracketed by CUDA events on the handle's compute stream; *ms = device time */
int mglc_lbm_step_timed(mglc_lbm *h, int nsteps, float *ms);
/* number of kernels this handle has launched so far / device time of the fused kernels only */
int mglc_lbm_launch_count(mglc_lbm *h, long long *n);
/* per-launch CUDA-event timing of the fused kernel on its launching stream (off by default);
 * mglc_lbm_kernel_time returns and resets the accumulated device time and launch count */
int mglc_lbm_set_profiling(mglc_lbm *h, int on);
/* How the fused step moves halos between subdomains (mode):
 *   2  direct halo stores: the fused kernel writes the outgoing populations of its boundary cells straight into the
 *      neighbours' halo cells over NVLink (peer access inside one process, CUDA IPC mappings between processes, set up
 *      collectively by mglc_lbm_create / mglc_group_create), followed by a flag barrier between neighbours.  No pack,
 *      send/recv or unpack; the transfer overlaps the update cell by cell.  Default when the mappings exist and the block has
 *      fewer than 2^25 cells (env MGLC_HALO_MODE=2|3 overrides the choice between 2 and 3).
 *   1  overlapped exchange: boundary shell first, pack -> ncclSend/ncclRecv -> unpack on a second stream beside the
 *      interior update, the schedule of collision_with_message_exchange, lid3_mpi_nonblock.f90:1108-1230.
 *      Default otherwise.
 *   0  blocking exchange, then update: message_passing_sendrecv() as the blocking driver does it, L3/main.f90:89-93.
 *   3  halo push: the plain fused kernel, then ONE small launch that copies every outgoing message (same sets) from the
 *      boundary cells straight into the neighbours' halo cells over the same mappings as mode 2 -- no send buffers, no NCCL,
 *      no unpack, and no message code inside the update kernel.  Default when the mappings exist and the block has 2^25 cells
 *      or more (384^3 and up: the stores of a face cost less from a separate launch than from inside the update).
 * With modes 2 and 3 every call that leaves the fused loop (upload, download, the per-subroutine entry points) must be made
 * by all ranks between the same two mglc_lbm_step calls, like the reference's subroutines; a rank out of step is
 * reported by mglc_lbm_sync / mglc_check as MGLC_E_STATE instead of hanging. */
int mglc_lbm_set_overlap(mglc_lbm *h, int mode);
int mglc_lbm_get_overlap(mglc_lbm *h, int *mode);   /* the transport in use (the library picks 3 for large blocks, 2 for small ones) */
/* 1 if the direct-halo mappings of mode 2 are established for this handle */
int mglc_lbm_direct_halo(mglc_lbm *h, int *available);
int mglc_lbm_kernel_time(mglc_lbm *h, float *fused_ms, long long *fused_launches);
/* page-locked host memory for the caller's f / rho,u,v,w arrays (c_f_pointer on the Fortran side), so
 * upload/download run at full PCIe rate; pageable pointers are accepted everywhere too */
int mglc_host_alloc(void **p, size_t bytes);
int mglc_host_free(void *p);
int mglc_lbm_device_bytes(mglc_lbm *h, long long *bytes);
int mglc_lbm_get_desc(mglc_lbm *h, mglc_lbm_desc *d);

/* ================= P subdomains in one process (harness / tests / single-process multi-GPU) ===== */
/* creates nranks subdomains of the global lattice gn with the reference decomposition; subdomain r
 * lives on devices[r] (NULL = all on device 0); halos move by device-to-device copies */
int mglc_group_create(mglc_group **g, const mglc_lbm_desc *global_desc, int nranks, const int *devices_or_null);
int mglc_group_destroy(mglc_group *g);
int mglc_group_size(mglc_group *g, int *nranks);
int mglc_group_rank(mglc_group *g, int r, mglc_lbm **h);      /* borrowed handle */
int mglc_group_initial(mglc_group *g);
int mglc_group_collision(mglc_group *g);
int mglc_group_exchange(mglc_group *g);
int mglc_group_streaming(mglc_group *g);
int mglc_group_bounceback(mglc_group *g);
int mglc_group_macro(mglc_group *g);
int mglc_group_check(mglc_group *g, double *errorU);
int mglc_group_collisionT(mglc_group *g);
int mglc_group_exchange_g(mglc_group *g);
int mglc_group_streamingT(mglc_group *g);
int mglc_group_bouncebackT(mglc_group *g);
int mglc_group_macroT(mglc_group *g);
int mglc_group_check_thermal(mglc_group *g, double *errorU, double *errorT);
int mglc_group_calNuRe(mglc_group *g, double prandtl, double *NuVolAvg, double *ReVolAvg);
int mglc_group_step(mglc_group *g, int nsteps);
int mglc_group_step_timed(mglc_group *g, int nsteps, float *ms);

/* ================= Jacobi halo-exchange path (MPI/Laplace/fortran/jacobi2d_mpi.f90, "LAP") =================
 * ndim = 2 is the reference's problem; ndim = 3 its six-neighbour extension (BASELINE.json config 2).
 * Host arrays are the reference's: column-major A(0:nx+1, 0:ny+1 [, 0:nz+1]) with one ghost layer
 * (LAP:78-81).  A handle owns either ONE subdomain (one process per GPU, halos over NCCL) or all P
 * subdomains of the process grid (mglc_jacobi_create_local: halos by device-to-device copies); `r` below
 * is the index among the subdomains the handle owns (0 for a one-subdomain handle). */
typedef struct mglc_jacobi mglc_jacobi;
/* MPI_Dims_create(nranks, ndim, dims) with dims = 0 -- LAP:47 (ndim = 2) / L3/main.f90:24 (ndim = 3) */
int mglc_dims_create_nd(int nranks, int ndim, int dims[3]);
/* MPI_Cart_create + decompose_1d + MPI_Cart_shift + allocate -- LAP:47-81 */
int mglc_jacobi_create(mglc_jacobi **h, int ndim, const int gn[3], const int dims_or_zero[3], int nranks,
                       int rank, int device, mglc_comm *comm_or_null);
int mglc_jacobi_create_local(mglc_jacobi **h, int ndim, const int gn[3], const int dims_or_zero[3],
                             int nranks, const int *devices_or_null);
int mglc_jacobi_destroy(mglc_jacobi *h);                                  /* deallocate -- LAP:121-124 */
int mglc_jacobi_nlocal(mglc_jacobi *h, int *n);
int mglc_jacobi_info(mglc_jacobi *h, int r, int dims[3], int ln[3], int start[3], int coords[3], int nbr[6]);
int mglc_jacobi_init(mglc_jacobi *h);                                     /* init() -- LAP:144-166 */
int mglc_jacobi_upload(mglc_jacobi *h, int r, const double *A, const double *A_new, const double *f); /* NULL = keep */
int mglc_jacobi_download(mglc_jacobi *h, int r, double *A, double *A_new);
int mglc_jacobi_exchange(mglc_jacobi *h);      /* exchange_message(A)              LAP:223-254 */
int mglc_jacobi_sweep(mglc_jacobi *h);         /* jacobi(A, A_new) + role swap     LAP:170-182, 97-103 */
int mglc_jacobi_step(mglc_jacobi *h, int nits);            /* nits x (exchange, sweep)  LAP:94-103 */
int mglc_jacobi_step_timed(mglc_jacobi *h, int nits, float *ms);
int mglc_jacobi_check_diff(mglc_jacobi *h, double *error_max);   /* check_diff + Allreduce(MAX)  LAP:105-107,185-204 */
int mglc_jacobi_launch_count(mglc_jacobi *h, long long *n);
int mglc_jacobi_sync(mglc_jacobi *h);
/* Halo transport of mglc_jacobi_step on several subdomains.  1 (default when the neighbours' arrays could be mapped: peer
 * pointers inside one process, CUDA IPC between processes): the sweep stores its boundary values straight into the
 * neighbours' ghost layers of A_new -- the values exchange_message() (LAP:223-254) would bring before the next sweep -- so the
 * transfer rides along with the sweep (the reference's own overlap of exchange and interior, jacobi2d_mpi_nonblock.f90:167-218,
 * taken to its end).  0: exchange_message, then sweep, as LAP:94-103 is written.  After mglc_jacobi_step in mode 1 the ghost
 * layers of A are already the neighbours' new boundary values; interior cells are identical in both modes. */
int mglc_jacobi_set_halo(mglc_jacobi *h, int mode);
int mglc_jacobi_direct_halo(mglc_jacobi *h, int *available);

/* ================= particle-laden D2Q9 path (MPI/Micro_particles/fortran/case4/mpi_particle/, "P4") =================
 * Host arrays are the reference's (P4/freeall.F90:13-17): f(0:8,-2:nx+3,-2:ny+3), f_post(0:8,-1:nx+2,-1:ny+2),
 * obst(0:nx+1,0:ny+1) integer, rho,u,v(nx,ny); column-major.  Particle state is replicated (size nparticles), like the
 * reference's module arrays.  The reference draws its initial particle positions from a compiler-specific
 * random_number (P4/initial.F90:52-75), so positions are an input here.  A handle owns ONE subdomain (one process
 * per GPU, halos and reductions over NCCL) or all P of them (mglc_p2d_create_local); `r` = index among those owned. */
typedef struct mglc_p2d mglc_p2d;
typedef struct mglc_p2d_desc {           /* module commondata, P4/commondata.F90:3-62 */
    int total_nx, total_ny;              /* 201 x 801                                      */
    int nparticles;                      /* cNumMax                                        */
    int reserved;
    double rho0, rhoSolid, viscosity;    /* 1, 1.01, 0.05 (tauf = 3 nu + 0.5)              */
    double radius0;                      /* 10                                             */
    double gravity;                      /* 980 * t0^2 / l0                                */
    double thresholdWall, stiffWall, thresholdParticle, stiffParticle;
    /* options taken from the reference's other particle scenario, MPI/Micro_particles/fortran/case1/mpi_complete; all 0 (as
     * mglc_p2d_desc_init leaves them) = case4 */
    double Uwall;                        /* top wall moves with +Uwall, bottom wall with -Uwall (case1/.../fluid.F90:123-171)   */
    double Uframe;                       /* U0 of the `movingFrame` build of that file; 0 = `stationaryFrame`                     */
    int bb_linear;                       /* 1: linear-interpolated bounce-back (case1/.../particle_bounceback.F90:66-76)         */
    int moving_walls;                    /* 1: apply the moving-wall terms                                                       */
} mglc_p2d_desc;
int mglc_p2d_desc_init(mglc_p2d_desc *d, int nparticles);            /* the shipped constants */
/* MPI_Dims_create_2d -- P4/mpi_starts.F90:160-180 (smallest halo message wins) */
int mglc_p2d_dims_create(int nranks, int total_nx, int total_ny, int dims[2]);
int mglc_p2d_create(mglc_p2d **h, const mglc_p2d_desc *d, const int dims_or_zero[2], int nranks, int rank, int device,
                    mglc_comm *comm_or_null);                          /* mpi_starts + allocate_all */
int mglc_p2d_create_local(mglc_p2d **h, const mglc_p2d_desc *d, const int dims_or_zero[2], int nranks,
                          const int *devices_or_null);
int mglc_p2d_destroy(mglc_p2d *h);                                      /* free_all -- P4/freeall.F90:23-41 */
int mglc_p2d_nlocal(mglc_p2d *h, int *n);
int mglc_p2d_info(mglc_p2d *h, int r, int dims[2], int ln[2], int start[2], int coords[2], int nbr[8]);
/* xCenter, yCenter, Uc, Vc, rationalOmega, radius; NULL = keep */
int mglc_p2d_set_particles(mglc_p2d *h, const double *x, const double *y, const double *U, const double *V,
                           const double *omega, const double *radius);
/* ... and wallTotalForceX/Y, totalTorque; NULL = skip */
int mglc_p2d_get_particles(mglc_p2d *h, double *x, double *y, double *U, double *V, double *omega, double *Fx,
                           double *Fy, double *torque);
int mglc_p2d_set_forces(mglc_p2d *h, const double *Fx, const double *Fy, const double *torque);
int mglc_p2d_upload(mglc_p2d *h, int r, const double *f, const double *f_post, const double *rho, const double *u,
                    const double *v, const int *obst);                  /* NULL = keep */
int mglc_p2d_download(mglc_p2d *h, int r, double *f, double *f_post, double *rho, double *u, double *v, int *obst);
int mglc_p2d_initial(mglc_p2d *h);               /* initial() after the positions are set   P4/initial.F90:79-199 */
int mglc_p2d_collision(mglc_p2d *h);             /* collision()            P4/fluid.F90:1-72                    */
int mglc_p2d_send_all_fp(mglc_p2d *h);           /* send_all_fp()          P4/message_send_all.F90:84-148       */
int mglc_p2d_streaming(mglc_p2d *h);             /* streaming()            P4/fluid.F90:75-111                  */
int mglc_p2d_bounceback(mglc_p2d *h);            /* bounceback()           P4/fluid.F90:115-161                 */
int mglc_p2d_bounceback_particle(mglc_p2d *h, int recompute_rho_avg);   /* P4/particle_bounceback.F90:1-96 (1 = as the reference) */
int mglc_p2d_macro(mglc_p2d *h);                 /* macro()                P4/fluid.F90:164-184                 */
int mglc_p2d_calforce(mglc_p2d *h);              /* calForce()             P4/particle_force.F90:1-212          */
int mglc_p2d_send_all_f(mglc_p2d *h);            /* send_all_f()           P4/message_send_all.F90:1-81         */
int mglc_p2d_update_center(mglc_p2d *h);         /* updateCenter()         P4/particle_update.F90:1-209         */
int mglc_p2d_check(mglc_p2d *h, double *errorU); /* check()                P4/fluid.F90:187-221                 */
int mglc_p2d_step(mglc_p2d *h, int nsteps);      /* nsteps loop bodies     P4/main.F90:35-73                    */
int mglc_p2d_step_timed(mglc_p2d *h, int nsteps, float *ms);
int mglc_p2d_set_rho_avg(mglc_p2d *h, double rhoAvg);
int mglc_p2d_get_rho_avg(mglc_p2d *h, double *rhoAvg);
/* the reference's stop / MPI_Abort conditions, raised on the device: returns MGLC_E_DIVERGED when any is set */
int mglc_p2d_error_flags(mglc_p2d *h, int *flags);
int mglc_p2d_launch_count(mglc_p2d *h, long long *n);
int mglc_p2d_sync(mglc_p2d *h);

/* ================= 2-D D2Q9 lid-driven cavity =================
 * The reference ships it as L2C = MPI/Lid_driven_cavity/c/lid_driven_cavity.c (plain C, one domain, 200 x 200) and
 * L2F = MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/ (Fortran + MPI, 2-D Cartesian blocks, 201 x 201).  They differ only
 * in the rounding of collision() (L2C: (...)/36.0 and meq(8) = u*v, c:186-255; L2F: per-term divisions and meq(8) = rho*u*v,
 * evolution.f90:15-70) and in check(); `variant` picks the one to reproduce.  A third program,
 * L2I = MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90 (sequential, 257 x 257), is L2F with the
 * incompressible equilibrium: meq without the rho factors (L2I:195-202), u, v = the undivided momentum sums (L2I:307-308), the lid
 * term -(+-U0)/6 without rho (L2I:291-292), initial() leaving rho = 0 and f = omega*(...) (L2I:137,160), check() = ratio of the
 * sums of dsqrt (L2I:327-335); MGLC_L2D_INCOMP reproduces it and decomposes like L2F.  MGLC_L2D_C_SRT is L2C with its own model
 * switch set to SRT (c:13-14): the single-relaxation-time collision c:160-176, the rest as L2C.  Host arrays are L2F's, column-major:
 * f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), rho,u,v(nx,ny) (initial.f90:30-38); L2C's f[NX][NY][9] is the same memory read as
 * (0:8,ny,nx), i.e. the caller transposes x and y.  A handle owns ONE subdomain (one process per GPU, halos over NCCL) or
 * all P of them (mglc_l2d_create_local); `r` = index among those owned. */
enum { MGLC_L2D_C = 0, MGLC_L2D_F = 1, MGLC_L2D_INCOMP = 2, MGLC_L2D_C_SRT = 3 };
typedef struct mglc_l2d mglc_l2d;
typedef struct mglc_l2d_desc {
    int total_nx, total_ny;              /* commondata.f90:4 ; c:9-10                       */
    int variant;                         /* MGLC_L2D_C | MGLC_L2D_F | MGLC_L2D_INCOMP | MGLC_L2D_C_SRT */
    int arith;                           /* MGLC_ARITH_FAST | MGLC_ARITH_STRICT             */
    double reynolds, U0, rho0;           /* 1000, 0.1, 1    commondata.f90:6-8 ; c:15-17    */
} mglc_l2d_desc;
int mglc_l2d_desc_init(mglc_l2d_desc *d, int variant);                  /* the shipped constants of that program */
/* MPI_Dims_create(np,2) + MPI_Cart_create + decompose_1d + MPI_Cart_shift + MPI_Cart_find_corners + allocate -- main.f90:24-60,
 * initial.f90:30-38; tau = U0*total_nx/Re*3 + 0.5, Snu, Sq -- commondata.f90:9,31 */
int mglc_l2d_create(mglc_l2d **h, const mglc_l2d_desc *d, const int dims_or_zero[2], int nranks, int rank, int device,
                    mglc_comm *comm_or_null);
int mglc_l2d_create_local(mglc_l2d **h, const mglc_l2d_desc *d, const int dims_or_zero[2], int nranks, const int *devices_or_null);
int mglc_l2d_destroy(mglc_l2d *h);
int mglc_l2d_nlocal(mglc_l2d *h, int *n);
/* nbr[0..3] = right(+x), left(-x), top(+y), bottom(-y); nbr[4..7] = where populations 5..8 go; -1 = MPI_PROC_NULL */
int mglc_l2d_info(mglc_l2d *h, int r, int dims[2], int ln[2], int start[2], int coords[2], int nbr[8]);
int mglc_l2d_params(mglc_l2d *h, double *tau, double *Snu, double *Sq);
int mglc_l2d_upload(mglc_l2d *h, int r, const double *f, const double *f_post, const double *rho, const double *u,
                    const double *v);                                   /* NULL = keep */
int mglc_l2d_download(mglc_l2d *h, int r, double *f, double *f_post, double *rho, double *u, double *v);
int mglc_l2d_initial(mglc_l2d *h);      /* initial()                    initial.f90:40-66   == c:123-151 */
int mglc_l2d_collision(mglc_l2d *h);    /* collision()                  evolution.f90:1-71   == c:186-255 */
int mglc_l2d_exchange(mglc_l2d *h);     /* message_passing_sendrecv()   ex_sendrecv.f90:1-81             */
int mglc_l2d_streaming(mglc_l2d *h);    /* streaming()                  evolution.f90:75-94  == c:262-283 */
int mglc_l2d_bounceback(mglc_l2d *h);   /* bounceback() / boundary()    bounceback.f90:1-44 == c:286-313 */
int mglc_l2d_macro(mglc_l2d *h);        /* macro()                      evolution.f90:98-112 == c:318-338 */
int mglc_l2d_check(mglc_l2d *h, double *errorU);   /* check()           evolution.f90:115-150 == c:341-363 */
int mglc_l2d_step(mglc_l2d *h, int nsteps);        /* nsteps loop bodies  main.f90:66-82    == c:57-63   */
int mglc_l2d_step_timed(mglc_l2d *h, int nsteps, float *ms);
int mglc_l2d_launch_count(mglc_l2d *h, long long *n);
int mglc_l2d_sync(mglc_l2d *h);

/* ================= 2-D thermal D2Q9 + D2Q5 driver =================
 * B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/ : D2Q9 MRT flow with the Boussinesq force Fy = rho*gBeta*(T-Tref),
 * D2Q5 MRT temperature; side-heated cell (macros.F90:24-27, shipped) or Rayleigh-Benard cell (macros.F90:17-20).
 * Arrays cross the boundary in the Fortran program's layout: f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), g(0:4,nx,ny),
 * g_post(0:4,0:nx+1,0:ny+1), rho,u,v,T,Fx,Fy(nx,ny) (initial.F90:177-197).  Handles own one subdomain (mglc_t2d_create) or all
 * P of them (mglc_t2d_create_local); `r` = index among those owned. */
typedef struct mglc_t2d mglc_t2d;
typedef struct mglc_t2d_desc {
    int total_nx, total_ny;              /* module.F90:26                                    */
    int arith;                           /* MGLC_ARITH_FAST | MGLC_ARITH_STRICT              */
    int bcT[4];                          /* +x (right), -x (left), +y (top), -y (bottom): MGLC_BCT_*   macros.F90:16-27;
                                            MGLC_BCT_PERIODIC on both vertical walls = VerticalWallsPeriodicalU + ...T of the
                                            OpenACC program (seq/bouyancy2d_acc.F90:15,22); needs dims[0] == 1 */
    int variant;                         /* MGLC_T2D_MPI: the mpi_blocked files | MGLC_T2D_ACC: seq/bouyancy2d_acc.F90 (its collision()
                                            rounds f_post(0) term by term, acc:679) */
    double Rayleigh, Prandtl, Mach;      /* module.F90:31-33                                 */
    double Thot, Tcold, Tref, rho0;      /* module.F90:67-68                                 */
    double lengthUnit;                   /* 0 = dble(total_ny) (module.F90:29); the OpenACC program uses dble(nx) (acc:57) */
    /* the sheared Rayleigh-Benard programs (seq/R_B_2d.F90, seq/bouyancy2d_omp.F90; 0 = the shipped MPI program):
     * Uwall = UwallTopLeft, TopRight, BottomLeft, BottomRight, LeftTop, LeftBottom, RightTop, RightBottom (R_B_2d.F90:118-120;
     * the halves meet at nxHalf / nyHalf, :87): initial() sets u, v on the walls (:466-481) and bounceback() subtracts
     * rho*C/6 from the diagonal populations off a wall (:790-898); cornersT = 1: its bouncebackT() corner cells (:1086-1106) */
    double Uwall[8];
    int cornersT;
} mglc_t2d_desc;
enum { MGLC_T2D_MPI = 0, MGLC_T2D_ACC = 1 };
int mglc_t2d_desc_init(mglc_t2d_desc *d);                               /* the shipped constants (201 x 201, Ra 1e7, side-heated) */
/* the OpenACC program as shipped: 513 x 257, Ra 1e5, Rayleigh-Benard plates, periodic vertical walls, lengthUnit = 513 (acc:9-22,55-60) */
int mglc_t2d_desc_init_acc(mglc_t2d_desc *d);
/* seq/R_B_2d.F90 as shipped: 201 x 201, Ra 1e7, Pr 5.3, Rayleigh-Benard plates, walls moving at shearReynolds = 100 (U0 =
 * 100*viscosity/ny; change total_ny / Prandtl / Rayleigh / Mach BEFORE calling, or recompute Uwall), cornersT = 1 (:56-61,79,118-120) */
int mglc_t2d_desc_init_sheared_rb(mglc_t2d_desc *d);
/* MPI_Dims_create(np,2) + MPI_Cart_create + decompose_1d + MPI_Cart_shift + MPI_Cart_find_corners + allocate -- main.F90:21-43,
 * initial.F90:177-197; tauf, viscosity, diffusivity, paraA, gBeta, Snu, Sq, Qd, Qnu -- module.F90:69-81; fails with
 * MGLC_E_INVALID where the reference stops (paraA outside (-4,1), initial.F90:30) */
int mglc_t2d_create(mglc_t2d **h, const mglc_t2d_desc *d, const int dims_or_zero[2], int nranks, int rank, int device,
                    mglc_comm *comm_or_null);
int mglc_t2d_create_local(mglc_t2d **h, const mglc_t2d_desc *d, const int dims_or_zero[2], int nranks, const int *devices_or_null);
int mglc_t2d_destroy(mglc_t2d *h);
int mglc_t2d_nlocal(mglc_t2d *h, int *n);
/* nbr[0..3] = right(+x), left(-x), top(+y), bottom(-y); nbr[4..7] = where populations 5..8 go; -1 = MPI_PROC_NULL */
int mglc_t2d_info(mglc_t2d *h, int r, int dims[2], int ln[2], int start[2], int coords[2], int nbr[8]);
/* out = tauf, viscosity, diffusivity, paraA, gBeta, Snu, Sq, Qd, Qnu, lengthUnit */
int mglc_t2d_params(mglc_t2d *h, double out[10]);
/* fields = rho, u, v, T, Fx, Fy; NULL (array or entry) = keep / skip */
int mglc_t2d_upload(mglc_t2d *h, int r, const double *f, const double *f_post, const double *g, const double *g_post,
                    const double *const fields_or_null[6]);
int mglc_t2d_download(mglc_t2d *h, int r, double *f, double *f_post, double *g, double *g_post, double *const fields_or_null[6]);
int mglc_t2d_initial(mglc_t2d *h);      /* initial()            initial.F90:199-335        */
int mglc_t2d_collision(mglc_t2d *h);    /* collision()          evolution_f.F90:1-84       */
int mglc_t2d_exchange_f(mglc_t2d *h);   /* message_passing_f()  message_exchange.F90:1-79   */
int mglc_t2d_streaming(mglc_t2d *h);    /* streaming()          evolution_f.F90:89-108     */
int mglc_t2d_bounceback(mglc_t2d *h);   /* bounceback()         evolution_f.F90:113-324    */
int mglc_t2d_collisionT(mglc_t2d *h);   /* collisionT()         evolution_g.F90:1-46       */
int mglc_t2d_exchange_g(mglc_t2d *h);   /* message_passing_g()  message_exchange.F90:85-118 */
int mglc_t2d_streamingT(mglc_t2d *h);   /* streamingT()         evolution_g.F90:49-68      */
int mglc_t2d_bouncebackT(mglc_t2d *h);  /* bouncebackT()        evolution_g.F90:71-159     */
int mglc_t2d_macro(mglc_t2d *h);        /* macro()              evolution_f.F90:328-342    */
int mglc_t2d_macroT(mglc_t2d *h);       /* macroT()             evolution_g.F90:163-176    */
int mglc_t2d_check(mglc_t2d *h, double *errorU, double *errorT);   /* check()   check.F90:1-51 */
/* calNuRe()'s volume averages: out = angular momentum / N, NuVolAvg, ReVolAvg -- NuRe.F90:27-78 */
int mglc_t2d_nure(mglc_t2d *h, double out[3]);
int mglc_t2d_step(mglc_t2d *h, int nsteps);                        /* nsteps loop bodies  main.F90:84-108 */
int mglc_t2d_step_timed(mglc_t2d *h, int nsteps, float *ms);
int mglc_t2d_launch_count(mglc_t2d *h, long long *n);
int mglc_t2d_sync(mglc_t2d *h);

/* ================= D3Q19 lid-driven cavity on ONE lattice: AA-pattern storage =================
 * The loop body of L3/main.f90:89-97 with the arithmetic of mglc_lbm_step (bit-identical to it in MGLC_ARITH_STRICT), updated in
 * place: 19 x 8 B of lattice per cell instead of 2 x 19 x 8 B, for the largest lattice one GPU can hold (about 1.4x the cells).
 * mglc_aa_create: one subdomain (all six faces are walls of the global box).  mglc_aa_group_create: the blocks of
 * mpi_starts/decompose_1d (L3/main.f90:33-63, 247-263) inside one process, on one or several devices that can address each
 * other: no halo message is packed, every launch stores what the neighbouring blocks need straight into their lattices and
 * the launches of neighbours are ordered by events (one neighbour barrier per launch).  The lattice
 * is only meaningful through these calls (between two streaming steps its populations sit in the opposite slots), so there are
 * no per-subroutine entry points: f and the fields are uploaded, stepped, checked and downloaded. */
typedef struct mglc_aa mglc_aa;
typedef struct mglc_aa_desc {
    int n[3];                            /* nx, ny, nz                                        L3/commondata.f90:4 */
    int arith;                           /* MGLC_ARITH_FAST | MGLC_ARITH_STRICT                                  */
    int collision;                       /* MGLC_MRT_LID | MGLC_BGK                           L3/collision.f90   */
    int device;
    double tau, U0, rho0;                /* tauf = U0*total_nx/Re*3 + 0.5                     L3/commondata.f90:6-9 */
} mglc_aa_desc;
int mglc_aa_desc_init(mglc_aa_desc *d, int nx, int ny, int nz, double reynolds, double U0, double rho0);
int mglc_aa_create(mglc_aa **h, const mglc_aa_desc *d);
int mglc_aa_destroy(mglc_aa *h);
int mglc_aa_initial(mglc_aa *h);                                        /* initial()   L3/initial.f90:55-73 */
/* f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz); NULL = keep */
int mglc_aa_upload(mglc_aa *h, const double *f, const double *rho, const double *u, const double *v, const double *w);
int mglc_aa_step(mglc_aa *h, int nsteps);                               /* nsteps loop bodies  L3/main.f90:89-97 */
int mglc_aa_step_timed(mglc_aa *h, int nsteps, float *ms);
int mglc_aa_check(mglc_aa *h, double *errorU);                          /* check()     L3/check.f90:12-34 */
int mglc_aa_download_macro(mglc_aa *h, double *rho, double *u, double *v, double *w);
int mglc_aa_download_f(mglc_aa *h, double *f);                          /* f as the reference holds it after the last loop body */
int mglc_aa_device_bytes(mglc_aa *h, long long *bytes);
int mglc_aa_launch_count(mglc_aa *h, long long *n);
int mglc_aa_sync(mglc_aa *h);
int mglc_aa_get_block(mglc_aa *h, int *ln, int *start);                 /* local size, 0-based global offset of a block */
/* decomposed: gd->n = the GLOBAL lattice (total_nx, total_ny, total_nz), tau from mglc_aa_desc_init with the global nx.
 * dims = {0,..} or NULL: MPI_Dims_create's factors (mglc_dims_create); devices = NULL: gd->device for every block.
 * Blocks are numbered like MPI_Cart_create's ranks (row-major, mglc_cart_rank).  On a block handle (mglc_aa_group_rank) the
 * per-block calls are mglc_aa_upload (NATURAL layout only), mglc_aa_download_macro / _download_f, mglc_aa_get_block,
 * mglc_aa_device_bytes, mglc_aa_launch_count; initial / step / check go through the group (else MGLC_E_STATE). */
typedef struct mglc_aa_group mglc_aa_group;
int mglc_aa_group_create(mglc_aa_group **g, const mglc_aa_desc *gd, int nranks, const int *dims, const int *devices);
int mglc_aa_group_destroy(mglc_aa_group *g);
int mglc_aa_group_size(mglc_aa_group *g, int *n);
int mglc_aa_group_dims(mglc_aa_group *g, int *dims);
int mglc_aa_group_rank(mglc_aa_group *g, int r, mglc_aa **h);
int mglc_aa_group_initial(mglc_aa_group *g);
int mglc_aa_group_step(mglc_aa_group *g, int nsteps);
int mglc_aa_group_step_timed(mglc_aa_group *g, int nsteps, float *ms);   /* longest block stream, ms */
int mglc_aa_group_check(mglc_aa_group *g, double *errorU);
int mglc_aa_group_sync(mglc_aa_group *g);
/* decomposed, one process per block (collective over the communicator; gd->n = the GLOBAL lattice).  The handle is this rank's
 * block: mglc_aa_initial / _step / _step_timed / _check (all-reduced) / _upload / _download_* are then collective calls made by
 * every rank, as the reference's subroutines are.  The neighbours' lattices are mapped through CUDA IPC; where that is not
 * possible (no NVLink / peer access) creation fails with MGLC_E_STATE: this path has no message transport. */
int mglc_aa_create_comm(mglc_aa **h, const mglc_aa_desc *gd, mglc_comm *comm, const int *dims);

/* ================= on-disk formats of the drivers' output()/backupData() (host-only) =================
 * All arrays are the reference's global (gathered) arrays, column-major (nx,ny,nz) -- what mglc_lbm_download_macro /
 * the Python gather hand back.  Unformatted files use the gfortran record framing the reference's Makefiles produce:
 * [int32 n][payload][int32 n], records above 2147483639 bytes split into subrecords (negative markers). */
enum { MGLC_FILE_LID_PLT = 0, MGLC_FILE_LID_BIN = 1, MGLC_FILE_LID_DAT = 2, MGLC_FILE_THERMAL_PLT = 3,
       MGLC_FILE_THERMAL_BIN = 4, MGLC_FILE_BACKUP = 5 };
/* one Fortran `write(unit) list` per record; max_subrecord_or_0 = 0 keeps gfortran's 2147483639 */
int mglc_unformatted_write(const char *path, int nrecords, const void *const *records, const long long *bytes,
                           long long max_subrecord_or_0);
int mglc_unformatted_read(const char *path, int nrecords, void *const *records, const long long *bytes);
/* output_binary(): records u, v, rho -- L3/output.f90:350-367 (w is not written by the reference) */
int mglc_output_binary_lid(const char *path, const double *u, const double *v, const double *rho, int nx, int ny, int nz);
/* output_binary(): records u, v, w, T -- B3:1593-1618 */
int mglc_output_binary_thermal(const char *path, const double *u, const double *v, const double *w, const double *T,
                               int nx, int ny, int nz);
/* backupData(): records u, v, w, T, f(0:18,nx,ny,nz), g(0:6,nx,ny,nz) -- B3/seq/bouyancy3d.F90:1011-1029;
 * mglc_backup_read = initial() with loadInitField = 1, seq:367-378 */
int mglc_backup_write(const char *path, const double *u, const double *v, const double *w, const double *T,
                      const double *f, const double *g, int nx, int ny, int nz);
int mglc_backup_read(const char *path, double *u, double *v, double *w, double *T, double *f, double *g, int nx, int ny, int nz);
/* the 2-D thermal driver: output_binary() records u, v, T -- Buoyancy_driven_cavity/fortran/2d/mpi_blocked/output.F90:192-217;
 * backupData() records f, g, u, v, T as the driver stores them -- output.F90:381-401 (f(0:8,nx,ny)), seq/bouyancy2d_acc.F90:1158-1189
 * (f(nx,ny,0:8)); mglc_backup_read_2d = initial() with loadInitField = 1 */
int mglc_output_binary_thermal2d(const char *path, const double *u, const double *v, const double *T, int nx, int ny);
int mglc_backup_write_2d(const char *path, const double *f, const double *g, const double *u, const double *v, const double *T,
                         int nx, int ny);
int mglc_backup_read_2d(const char *path, double *f, double *g, double *u, double *v, double *T, int nx, int ny);
/* xp(0:total_n+1): 0, i - 0.5, total_n -- L3/initial.f90:18-31, B3:459-472 */
int mglc_grid_coords(int total_n, double *xp);
/* output_Tecplot(): "#!TDV101", one ordered zone, POINT packing, 7 float variables X Y Z U V W Pressure(= rho/3)
 * -- L3/output.f90:175-313; thermal: ... W T -- B3:1623-1773.  xp,yp,zp are (0:n+1) as above */
int mglc_output_tecplot_lid(const char *path, const double *xp, const double *yp, const double *zp, const double *u,
                            const double *v, const double *w, const double *rho, int nx, int ny, int nz);
int mglc_output_tecplot_thermal(const char *path, const double *xp, const double *yp, const double *zp, const double *u,
                                const double *v, const double *w, const double *T, int nx, int ny, int nz);
/* getVelocity(): the centre-line profiles u(nxHalf,nyHalf,:)/U0 vs zp/nz and w(:,nyHalf,nzHalf)/U0 vs xp/nx, as numbers
 * (the reference prints them list-directed, which is compiler-specific) -- L3/output.f90:318-347 */
int mglc_get_velocity(const double *xp, const double *zp, const double *u, const double *w, int nx, int ny, int nz,
                      double U0, double *uz, double *zn, double *xn, double *wx);
/* the drivers' file names for iteration itc: kind = MGLC_FILE_* */
int mglc_output_filename(char *out, size_t cap, int kind, int itc);

#ifdef __cplusplus
}
#endif
#endif /* MGLC_H */
