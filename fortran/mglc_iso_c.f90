!> ISO_C_BINDING interface to libmglc.so (include/mglc.h) for MGLC's Fortran drivers.
!!
!! The reference has no bind(C) layer; its time loops call argument-less subroutines over
!! `module commondata` (MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/main.f90:85-103).  With this
!! module a driver keeps its own `main`, arrays and output routines and swaps the loop body for the
!! GPU path, one call at a time or all at once (see INTEGRATION.md).
!!
!! Not compiled in the authoring image (no Fortran compiler there); it is declarative and mirrors the
!! C header one to one.  Link with:  mpif90 ... mglc_iso_c.f90 main.f90 -L<repo>/mglc_b200 -lmglc
module mglc_iso_c
    use, intrinsic :: iso_c_binding
    implicit none

    integer(c_int), parameter :: MGLC_OK = 0
    integer(c_int), parameter :: MGLC_D3Q19 = 0, MGLC_D3Q19_D3Q7 = 1, MGLC_MRT_LID = 0, MGLC_MRT_THERMAL = 1
    integer(c_int), parameter :: MGLC_BCT_ADIABATIC = 0, MGLC_BCT_CONST_HOT = 1, MGLC_BCT_CONST_COLD = 2
    integer(c_int), parameter :: MGLC_ARITH_FAST = 0, MGLC_ARITH_STRICT = 1

    !> mglc_lbm_desc: replaces the compile-time parameters of commondata (commondata.f90:4-15,42-53)
    type, bind(C) :: mglc_lbm_desc
        integer(c_int) :: lattice, collision, arith, kernel
        integer(c_int) :: gn(3), dims(3), coords(3), ln(3), start(3)
        real(c_double) :: tau, U0, rho0
        integer(c_int) :: device
        integer(c_int) :: bcT(6)                  ! thermal wall kinds, +x,-x,+y,-y,+z,-z (bouyancy3d_mpi.F90:5-19)
        integer(c_int) :: reserved(1)
        real(c_double) :: paraA, gBeta, Tref, Thot, Tcold, omegaRot, Qd, Qnu   ! bouyancy3d_mpi.F90:33-47,73-74
    end type mglc_lbm_desc

    !> mglc_p2d_desc: module commondata of the particle driver (case4/mpi_particle/commondata.F90:3-62)
    type, bind(C) :: mglc_p2d_desc
        integer(c_int) :: total_nx, total_ny, nparticles, reserved
        real(c_double) :: rho0, rhoSolid, viscosity, radius0, gravity
        real(c_double) :: thresholdWall, stiffWall, thresholdParticle, stiffParticle
        real(c_double) :: Uwall, Uframe            ! case1/mpi_complete options: moving walls (fluid.F90:123-171); 0 = case4
        integer(c_int) :: bb_linear, moving_walls  ! linear-interpolated bounce-back (particle_bounceback.F90:66-76)
    end type mglc_p2d_desc
    !> mglc_l2d_desc: module commondata of the 2-D lid driver (Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/commondata.f90:4-9)
    type, bind(C) :: mglc_l2d_desc
        integer(c_int) :: total_nx, total_ny, variant, arith
        real(c_double) :: reynolds, U0, rho0
    end type mglc_l2d_desc
    integer(c_int), parameter :: MGLC_L2D_C = 0, MGLC_L2D_F = 1, MGLC_L2D_INCOMP = 2, MGLC_L2D_C_SRT = 3
    !> mglc_aa_desc: the lid driver's constants (L3/commondata.f90:4-9) for the single-lattice (AA-pattern) path
    type, bind(C) :: mglc_aa_desc
        integer(c_int) :: n(3), arith, collision, device
        real(c_double) :: tau, U0, rho0
    end type mglc_aa_desc
    !> mglc_t2d_desc: module commondata of the 2-D thermal driver (Buoyancy_driven_cavity/fortran/2d/mpi_blocked/module.F90:26-33,67-68)
    !> and the boundary macro set of macros.F90:16-27 (bcT = +x, -x, +y, -y)
    type, bind(C) :: mglc_t2d_desc
        integer(c_int) :: total_nx, total_ny, arith, bcT(4), variant
        real(c_double) :: Rayleigh, Prandtl, Mach, Thot, Tcold, Tref, rho0, lengthUnit
        !> the sheared Rayleigh-Benard programs (seq/R_B_2d.F90:118-120, :1086-1106): UwallTopLeft, TopRight, BottomLeft,
        !> BottomRight, LeftTop, LeftBottom, RightTop, RightBottom; cornersT = 1: its bouncebackT() corner cells
        real(c_double) :: Uwall(8)
        integer(c_int) :: cornersT
    end type mglc_t2d_desc
    integer(c_int), parameter :: MGLC_T2D_MPI = 0, MGLC_T2D_ACC = 1, MGLC_BCT_PERIODIC = 3

    interface
        ! ---- host-only helpers -------------------------------------------------------------------
        function mglc_lbm_desc_init(d, gn, dims_or_zero, nranks, rank, reynolds, U0, rho0) &
                bind(C, name="mglc_lbm_desc_init") result(rc)
            import :: c_int, c_double, mglc_lbm_desc
            type(mglc_lbm_desc), intent(out) :: d
            integer(c_int), intent(in) :: gn(3), dims_or_zero(3)
            integer(c_int), value :: nranks, rank
            real(c_double), value :: reynolds, U0, rho0
            integer(c_int) :: rc
        end function
        function mglc_last_error() bind(C, name="mglc_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        ! ---- communicator: the id is MPI_Bcast by the driver --------------------------------------
        function mglc_comm_unique_id(id) bind(C, name="mglc_comm_unique_id") result(rc)
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
            integer(c_int) :: rc
        end function
        function mglc_comm_init_rank(comm, id, nranks, rank, device) bind(C, name="mglc_comm_init_rank") result(rc)
            import :: c_int, c_char, c_ptr
            type(c_ptr), intent(out) :: comm
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int), value :: nranks, rank, device
            integer(c_int) :: rc
        end function
        function mglc_comm_destroy(comm) bind(C, name="mglc_comm_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        ! ---- one subdomain ------------------------------------------------------------------------
        function mglc_lbm_create(h, d, comm) bind(C, name="mglc_lbm_create") result(rc)
            import :: c_int, c_ptr, mglc_lbm_desc
            type(c_ptr), intent(out) :: h
            type(mglc_lbm_desc), intent(in) :: d
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        function mglc_lbm_destroy(h) bind(C, name="mglc_lbm_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_lbm_initial(h) bind(C, name="mglc_lbm_initial") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        !> f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz): the driver's own arrays, passed as they are
        function mglc_lbm_upload(h, f, rho, u, v, w) bind(C, name="mglc_lbm_upload") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: f(*), rho(*), u(*), v(*), w(*)
            integer(c_int) :: rc
        end function
        function mglc_lbm_download_macro(h, rho, u, v, w) bind(C, name="mglc_lbm_download_macro") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: rho(*), u(*), v(*), w(*)
            integer(c_int) :: rc
        end function
        function mglc_lbm_download_f(h, f) bind(C, name="mglc_lbm_download_f") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: f(*)
            integer(c_int) :: rc
        end function
        ! one call per reference subroutine
        function mglc_collision(h) bind(C, name="mglc_collision") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_exchange(h) bind(C, name="mglc_exchange") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_streaming(h) bind(C, name="mglc_streaming") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_bounceback(h) bind(C, name="mglc_bounceback") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_macro(h) bind(C, name="mglc_macro") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_check(h, errorU) bind(C, name="mglc_check") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU
            integer(c_int) :: rc
        end function
        !> nsteps iterations of the loop body main.f90:85-97 in fused form
        function mglc_lbm_step(h, nsteps) bind(C, name="mglc_lbm_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nsteps
            integer(c_int) :: rc
        end function
        function mglc_lbm_sync(h) bind(C, name="mglc_lbm_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_host_alloc(p, bytes) bind(C, name="mglc_host_alloc") result(rc)
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), intent(out) :: p
            integer(c_size_t), value :: bytes
            integer(c_int) :: rc
        end function
        function mglc_host_free(p) bind(C, name="mglc_host_free") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: p
            integer(c_int) :: rc
        end function

        ! ---- thermal double-distribution driver (Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90:222-248) --------
        function mglc_thermal_desc_init(d, gn, dims_or_zero, nranks, rank, rayleigh, prandtl, mach, ekman) &
                bind(C, name="mglc_thermal_desc_init") result(rc)
            import :: c_int, c_double, mglc_lbm_desc
            type(mglc_lbm_desc), intent(out) :: d
            integer(c_int), intent(in) :: gn(3), dims_or_zero(3)
            integer(c_int), value :: nranks, rank
            real(c_double), value :: rayleigh, prandtl, mach, ekman
            integer(c_int) :: rc
        end function
        !> g(0:6,nx,ny,nz), T, Fx, Fy, Fz (nx,ny,nz)
        function mglc_lbm_upload_thermal(h, g, T, Fx, Fy, Fz) bind(C, name="mglc_lbm_upload_thermal") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: g(*), T(*), Fx(*), Fy(*), Fz(*)
            integer(c_int) :: rc
        end function
        function mglc_lbm_download_thermal(h, g, T, Fx, Fy, Fz) bind(C, name="mglc_lbm_download_thermal") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: g(*), T(*), Fx(*), Fy(*), Fz(*)
            integer(c_int) :: rc
        end function
        function mglc_collisionT(h) bind(C, name="mglc_collisionT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_exchange_g(h) bind(C, name="mglc_exchange_g") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_streamingT(h) bind(C, name="mglc_streamingT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_bouncebackT(h) bind(C, name="mglc_bouncebackT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_macroT(h) bind(C, name="mglc_macroT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_check_thermal(h, errorU, errorT) bind(C, name="mglc_check_thermal") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU, errorT
            integer(c_int) :: rc
        end function
        ! ---- Jacobi driver (MPI/Laplace/fortran/jacobi2d_mpi.f90:94-112) ---------------------------------------------
        function mglc_jacobi_create(h, ndim, gn, dims_or_zero, nranks, rank, device, comm) bind(C, name="mglc_jacobi_create") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: ndim, nranks, rank, device
            integer(c_int), intent(in) :: gn(3), dims_or_zero(3)
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        !> A(0:nx+1,0:ny+1[,0:nz+1]) exactly as allocated at jacobi2d_mpi.f90:78-81; c_null_ptr-free variants: pass the arrays
        function mglc_jacobi_upload(h, r, A, A_new, f) bind(C, name="mglc_jacobi_upload") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(in) :: A(*), A_new(*), f(*)
            integer(c_int) :: rc
        end function
        function mglc_jacobi_download(h, r, A, A_new) bind(C, name="mglc_jacobi_download") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(out) :: A(*), A_new(*)
            integer(c_int) :: rc
        end function
        function mglc_jacobi_step(h, nits) bind(C, name="mglc_jacobi_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nits
            integer(c_int) :: rc
        end function
        function mglc_jacobi_check_diff(h, error_max) bind(C, name="mglc_jacobi_check_diff") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: error_max
            integer(c_int) :: rc
        end function
        function mglc_jacobi_destroy(h) bind(C, name="mglc_jacobi_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- particle driver (Micro_particles/fortran/case4/mpi_particle/main.F90:35-73) -----------------------------
        function mglc_p2d_desc_init(d, nparticles) bind(C, name="mglc_p2d_desc_init") result(rc)
            import :: c_int, mglc_p2d_desc
            type(mglc_p2d_desc), intent(out) :: d
            integer(c_int), value :: nparticles
            integer(c_int) :: rc
        end function
        function mglc_p2d_create(h, d, dims_or_zero, nranks, rank, device, comm) bind(C, name="mglc_p2d_create") result(rc)
            import :: c_int, c_ptr, mglc_p2d_desc
            type(c_ptr), intent(out) :: h
            type(mglc_p2d_desc), intent(in) :: d
            integer(c_int), intent(in) :: dims_or_zero(2)
            integer(c_int), value :: nranks, rank, device
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        !> xCenter, yCenter, Uc, Vc, rationalOmega, radius (cNumMax each)
        function mglc_p2d_set_particles(h, x, y, U, V, omega, radius) bind(C, name="mglc_p2d_set_particles") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: x(*), y(*), U(*), V(*), omega(*), radius(*)
            integer(c_int) :: rc
        end function
        function mglc_p2d_get_particles(h, x, y, U, V, omega, Fx, Fy, torque) bind(C, name="mglc_p2d_get_particles") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: x(*), y(*), U(*), V(*), omega(*), Fx(*), Fy(*), torque(*)
            integer(c_int) :: rc
        end function
        function mglc_p2d_initial(h) bind(C, name="mglc_p2d_initial") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        !> nsteps x (collision, send_all_fp, streaming, bounceback, bounceback_particle, macro, calForce, send_all_f, updateCenter)
        function mglc_p2d_step(h, nsteps) bind(C, name="mglc_p2d_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nsteps
            integer(c_int) :: rc
        end function
        function mglc_p2d_check(h, errorU) bind(C, name="mglc_p2d_check") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU
            integer(c_int) :: rc
        end function
        !> the reference's stop / MPI_Abort conditions as a flag word; rc = MGLC_E_DIVERGED when any is set
        function mglc_p2d_error_flags(h, flags) bind(C, name="mglc_p2d_error_flags") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), intent(out) :: flags
            integer(c_int) :: rc
        end function
        function mglc_p2d_download(h, r, f, f_post, rho, u, v, obst) bind(C, name="mglc_p2d_download") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(out) :: f(*), f_post(*), rho(*), u(*), v(*)
            integer(c_int), intent(out) :: obst(*)
            integer(c_int) :: rc
        end function
        function mglc_p2d_destroy(h) bind(C, name="mglc_p2d_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- 2-D lid driver (Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/main.f90:66-82) ------------------------
        function mglc_l2d_desc_init(d, variant) bind(C, name="mglc_l2d_desc_init") result(rc)
            import :: c_int, mglc_l2d_desc
            type(mglc_l2d_desc), intent(out) :: d
            integer(c_int), value :: variant
            integer(c_int) :: rc
        end function
        function mglc_l2d_create(h, d, dims_or_zero, nranks, rank, device, comm) bind(C, name="mglc_l2d_create") result(rc)
            import :: c_int, c_ptr, mglc_l2d_desc
            type(c_ptr), intent(out) :: h
            type(mglc_l2d_desc), intent(in) :: d
            integer(c_int), intent(in) :: dims_or_zero(2)
            integer(c_int), value :: nranks, rank, device
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        function mglc_l2d_info(h, r, dims, ln, start, coords, nbr) bind(C, name="mglc_l2d_info") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: r
            integer(c_int), intent(out) :: dims(2), ln(2), start(2), coords(2), nbr(8)
            integer(c_int) :: rc
        end function
        !> f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), rho,u,v(nx,ny) exactly as allocated at initial.f90:30-38
        function mglc_l2d_upload(h, r, f, f_post, rho, u, v) bind(C, name="mglc_l2d_upload") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(in) :: f(*), f_post(*), rho(*), u(*), v(*)
            integer(c_int) :: rc
        end function
        function mglc_l2d_download(h, r, f, f_post, rho, u, v) bind(C, name="mglc_l2d_download") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(out) :: f(*), f_post(*), rho(*), u(*), v(*)
            integer(c_int) :: rc
        end function
        !> nsteps x (collision, message_passing_sendrecv, streaming, bounceback, macro)
        function mglc_l2d_step(h, nsteps) bind(C, name="mglc_l2d_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nsteps
            integer(c_int) :: rc
        end function
        function mglc_l2d_check(h, errorU) bind(C, name="mglc_l2d_check") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU
            integer(c_int) :: rc
        end function
        function mglc_l2d_destroy(h) bind(C, name="mglc_l2d_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- lid driver on ONE lattice (AA-pattern storage): L3/main.f90:89-97 for the largest single-GPU lattices -----------
        function mglc_aa_desc_init(d, nx, ny, nz, reynolds, U0, rho0) bind(C, name="mglc_aa_desc_init") result(rc)
            import :: c_int, c_double, mglc_aa_desc
            type(mglc_aa_desc), intent(out) :: d
            integer(c_int), value :: nx, ny, nz
            real(c_double), value :: reynolds, U0, rho0
            integer(c_int) :: rc
        end function
        function mglc_aa_create(h, d) bind(C, name="mglc_aa_create") result(rc)
            import :: c_int, c_ptr, mglc_aa_desc
            type(c_ptr), intent(out) :: h
            type(mglc_aa_desc), intent(in) :: d
            integer(c_int) :: rc
        end function
        !> f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz) as allocated at L3/initial.f90:35-44
        function mglc_aa_upload(h, f, rho, u, v, w) bind(C, name="mglc_aa_upload") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: f(*), rho(*), u(*), v(*), w(*)
            integer(c_int) :: rc
        end function
        function mglc_aa_step(h, nsteps) bind(C, name="mglc_aa_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nsteps
            integer(c_int) :: rc
        end function
        function mglc_aa_check(h, errorU) bind(C, name="mglc_aa_check") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU
            integer(c_int) :: rc
        end function
        function mglc_aa_download_macro(h, rho, u, v, w) bind(C, name="mglc_aa_download_macro") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: rho(*), u(*), v(*), w(*)
            integer(c_int) :: rc
        end function
        function mglc_aa_download_f(h, f) bind(C, name="mglc_aa_download_f") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: f(*)
            integer(c_int) :: rc
        end function
        function mglc_aa_destroy(h) bind(C, name="mglc_aa_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- 2-D thermal driver (Buoyancy_driven_cavity/fortran/2d/mpi_blocked/main.F90:84-108) ---------------------------
        function mglc_t2d_desc_init(d) bind(C, name="mglc_t2d_desc_init") result(rc)
            import :: c_int, mglc_t2d_desc
            type(mglc_t2d_desc), intent(out) :: d
            integer(c_int) :: rc
        end function
        !> seq/R_B_2d.F90 as shipped: Rayleigh-Benard plates, Pr = 5.3, walls moving at shearReynolds = 100
        function mglc_t2d_desc_init_sheared_rb(d) bind(C, name="mglc_t2d_desc_init_sheared_rb") result(rc)
            import :: c_int, mglc_t2d_desc
            type(mglc_t2d_desc), intent(out) :: d
            integer(c_int) :: rc
        end function
        !> the OpenACC program's shipped constants, seq/bouyancy2d_acc.F90:9-22,55-60
        function mglc_t2d_desc_init_acc(d) bind(C, name="mglc_t2d_desc_init_acc") result(rc)
            import :: c_int, mglc_t2d_desc
            type(mglc_t2d_desc), intent(out) :: d
            integer(c_int) :: rc
        end function
        function mglc_t2d_create(h, d, dims_or_zero, nranks, rank, device, comm) bind(C, name="mglc_t2d_create") result(rc)
            import :: c_int, c_ptr, mglc_t2d_desc
            type(c_ptr), intent(out) :: h
            type(mglc_t2d_desc), intent(in) :: d
            integer(c_int), intent(in) :: dims_or_zero(2)
            integer(c_int), value :: nranks, rank, device
            type(c_ptr), value :: comm
            integer(c_int) :: rc
        end function
        function mglc_t2d_info(h, r, dims, ln, start, coords, nbr) bind(C, name="mglc_t2d_info") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: r
            integer(c_int), intent(out) :: dims(2), ln(2), start(2), coords(2), nbr(8)
            integer(c_int) :: rc
        end function
        !> f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), g(0:4,nx,ny), g_post(0:4,0:nx+1,0:ny+1) as allocated at initial.F90:191-194;
        !> fields = c_loc of rho, u, v, T, Fx, Fy (nx,ny) in that order, c_null_ptr = keep
        function mglc_t2d_upload(h, r, f, f_post, g, g_post, fields) bind(C, name="mglc_t2d_upload") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(in) :: f(*), f_post(*), g(*), g_post(*)
            type(c_ptr), intent(in) :: fields(6)
            integer(c_int) :: rc
        end function
        function mglc_t2d_download(h, r, f, f_post, g, g_post, fields) bind(C, name="mglc_t2d_download") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: r
            real(c_double), intent(out) :: f(*), f_post(*), g(*), g_post(*)
            type(c_ptr), intent(in) :: fields(6)
            integer(c_int) :: rc
        end function
        !> nsteps x (collision, message_passing_f, streaming, bounceback, collisionT, message_passing_g, streamingT, bouncebackT, macro, macroT)
        function mglc_t2d_step(h, nsteps) bind(C, name="mglc_t2d_step") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), value :: nsteps
            integer(c_int) :: rc
        end function
        !> check(), check.F90:1-51
        function mglc_t2d_check(h, errorU, errorT) bind(C, name="mglc_t2d_check") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: errorU, errorT
            integer(c_int) :: rc
        end function
        !> calNuRe()'s volume averages (angular momentum / N, NuVolAvg, ReVolAvg), NuRe.F90:27-78
        function mglc_t2d_nure(h, out3) bind(C, name="mglc_t2d_nure") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: out3(3)
            integer(c_int) :: rc
        end function
        function mglc_t2d_destroy(h) bind(C, name="mglc_t2d_destroy") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- one call per reference subroutine (Laplace driver, jacobi2d_mpi.f90:94-112): handle in, status out ----
        function mglc_jacobi_init(h) bind(C, name="mglc_jacobi_init") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_jacobi_exchange(h) bind(C, name="mglc_jacobi_exchange") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_jacobi_sweep(h) bind(C, name="mglc_jacobi_sweep") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_jacobi_sync(h) bind(C, name="mglc_jacobi_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- one call per reference subroutine (particle driver, case4/mpi_particle/main.F90:37-71): handle in, status out ----
        function mglc_p2d_collision(h) bind(C, name="mglc_p2d_collision") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_send_all_fp(h) bind(C, name="mglc_p2d_send_all_fp") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_streaming(h) bind(C, name="mglc_p2d_streaming") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_bounceback(h) bind(C, name="mglc_p2d_bounceback") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_macro(h) bind(C, name="mglc_p2d_macro") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_calforce(h) bind(C, name="mglc_p2d_calforce") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_send_all_f(h) bind(C, name="mglc_p2d_send_all_f") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_update_center(h) bind(C, name="mglc_p2d_update_center") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_p2d_sync(h) bind(C, name="mglc_p2d_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- one call per reference subroutine (2-D lid driver, 2d_revised/mpi_blocked/main.f90:66-82): handle in, status out ----
        function mglc_l2d_initial(h) bind(C, name="mglc_l2d_initial") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_collision(h) bind(C, name="mglc_l2d_collision") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_exchange(h) bind(C, name="mglc_l2d_exchange") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_streaming(h) bind(C, name="mglc_l2d_streaming") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_bounceback(h) bind(C, name="mglc_l2d_bounceback") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_macro(h) bind(C, name="mglc_l2d_macro") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_l2d_sync(h) bind(C, name="mglc_l2d_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- one call per reference subroutine (2-D thermal driver, Buoyancy_driven_cavity/fortran/2d/mpi_blocked/main.F90:84-108): handle in, status out ----
        function mglc_t2d_initial(h) bind(C, name="mglc_t2d_initial") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_collision(h) bind(C, name="mglc_t2d_collision") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_exchange_f(h) bind(C, name="mglc_t2d_exchange_f") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_streaming(h) bind(C, name="mglc_t2d_streaming") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_bounceback(h) bind(C, name="mglc_t2d_bounceback") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_collisionT(h) bind(C, name="mglc_t2d_collisionT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_exchange_g(h) bind(C, name="mglc_t2d_exchange_g") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_streamingT(h) bind(C, name="mglc_t2d_streamingT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_bouncebackT(h) bind(C, name="mglc_t2d_bouncebackT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_macro(h) bind(C, name="mglc_t2d_macro") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_macroT(h) bind(C, name="mglc_t2d_macroT") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_t2d_sync(h) bind(C, name="mglc_t2d_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        ! ---- one call per reference subroutine (lid driver on one lattice): handle in, status out ----
        function mglc_aa_initial(h) bind(C, name="mglc_aa_initial") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        function mglc_aa_sync(h) bind(C, name="mglc_aa_sync") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int) :: rc
        end function
        !> the MPI lid driver on one lattice per rank: d describes the GLOBAL lattice (total_nx, total_ny, total_nz), comm as for
        !> mglc_lbm_create, dims = the process grid or zeros (MPI_Dims_create's).  mglc_aa_initial / _step / _check / _upload /
        !> _download_* on the handle are then collective calls, arrays are this rank's block (L3/main.f90:33-63)
        function mglc_aa_create_comm(h, d, comm, dims) bind(C, name="mglc_aa_create_comm") result(rc)
            import :: c_int, c_ptr, mglc_aa_desc
            type(c_ptr), intent(out) :: h
            type(mglc_aa_desc), intent(in) :: d
            type(c_ptr), value :: comm
            integer(c_int), intent(in) :: dims(3)
            integer(c_int) :: rc
        end function
        !> ln = nx, ny, nz of this rank's block; start = i_start_global - 1, j_start_global - 1, k_start_global - 1
        function mglc_aa_get_block(h, ln, start) bind(C, name="mglc_aa_get_block") result(rc)
            import :: c_int, c_ptr
            type(c_ptr), value :: h
            integer(c_int), intent(out) :: ln(3), start(3)
            integer(c_int) :: rc
        end function
        ! ---- in-loop diagnostics on the device and the drivers' on-disk formats ----------------------------------------
        !> calNuRe(), Buoyancy_driven_cavity/fortran/3d/mpi_blocked/RaNu.F90:13-47
        function mglc_calNuRe(h, prandtl, NuVolAvg, ReVolAvg) bind(C, name="mglc_calNuRe") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            real(c_double), value :: prandtl
            real(c_double), intent(out) :: NuVolAvg, ReVolAvg
            integer(c_int) :: rc
        end function
        !> one line of rho|u|v|w|T (field 0..4) along axis through the global 1-based (g1, g2), e.g. getVelocity(), L3/output.f90:334-344
        function mglc_lbm_download_line(h, field, axis, g1, g2, out, first, count) bind(C, name="mglc_lbm_download_line") result(rc)
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: h
            integer(c_int), value :: field, axis, g1, g2
            real(c_double), intent(out) :: out(*)
            integer(c_int), intent(out) :: first, count
            integer(c_int) :: rc
        end function
        !> backupData(), Buoyancy_driven_cavity/fortran/3d/seq/bouyancy3d.F90:1011-1029 (path is a C string: trim(name)//c_null_char)
        function mglc_backup_write(path, u, v, w, T, f, g, nx, ny, nz) bind(C, name="mglc_backup_write") result(rc)
            import :: c_int, c_char, c_double
            character(kind=c_char), intent(in) :: path(*)
            real(c_double), intent(in) :: u(*), v(*), w(*), T(*), f(*), g(*)
            integer(c_int), value :: nx, ny, nz
            integer(c_int) :: rc
        end function
        !> initial() with loadInitField = 1, seq/bouyancy3d.F90:367-378
        function mglc_backup_read(path, u, v, w, T, f, g, nx, ny, nz) bind(C, name="mglc_backup_read") result(rc)
            import :: c_int, c_char, c_double
            character(kind=c_char), intent(in) :: path(*)
            real(c_double), intent(out) :: u(*), v(*), w(*), T(*), f(*), g(*)
            integer(c_int), value :: nx, ny, nz
            integer(c_int) :: rc
        end function
        !> output_Tecplot(), L3/output.f90:175-313
        function mglc_output_tecplot_lid(path, xp, yp, zp, u, v, w, rho, nx, ny, nz) bind(C, name="mglc_output_tecplot_lid") result(rc)
            import :: c_int, c_char, c_double
            character(kind=c_char), intent(in) :: path(*)
            real(c_double), intent(in) :: xp(*), yp(*), zp(*), u(*), v(*), w(*), rho(*)
            integer(c_int), value :: nx, ny, nz
            integer(c_int) :: rc
        end function
        !> output_binary(), L3/output.f90:350-367
        function mglc_output_binary_lid(path, u, v, rho, nx, ny, nz) bind(C, name="mglc_output_binary_lid") result(rc)
            import :: c_int, c_char, c_double
            character(kind=c_char), intent(in) :: path(*)
            real(c_double), intent(in) :: u(*), v(*), rho(*)
            integer(c_int), value :: nx, ny, nz
            integer(c_int) :: rc
        end function
    end interface
end module mglc_iso_c
