"""ctypes binding of libmglc.so (the C ABI in include/mglc.h).

The library is the product; this module only declares its entry points.  There is no fallback: if the
shared object is missing the import fails, and without a CUDA device the device entry points return
MGLC_E_NOGPU, which `check()` turns into an exception.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmglc.so")

OK, E_INVALID, E_CUDA, E_NCCL, E_NOMEM, E_STATE, E_NOGPU, E_DIVERGED = 0, -1, -2, -3, -4, -5, -6, -7
D3Q19, D3Q19_D3Q7, D2Q9 = 0, 1, 2
MRT_LID, MRT_THERMAL, BGK = 0, 1, 2
ARITH_FAST, ARITH_STRICT = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA = 0, 1, 2
BCT_ADIABATIC, BCT_CONST_HOT, BCT_CONST_COLD, BCT_PERIODIC = 0, 1, 2, 3
T2D_MPI, T2D_ACC = 0, 1
L2D_C, L2D_F, L2D_INCOMP, L2D_C_SRT = 0, 1, 2, 3


class MglcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmglc error {code}: {msg}")
        self.code = code


class LbmDesc(C.Structure):
    _fields_ = [("lattice", C.c_int), ("collision", C.c_int), ("arith", C.c_int), ("kernel", C.c_int),
                ("gn", C.c_int * 3), ("dims", C.c_int * 3), ("coords", C.c_int * 3), ("ln", C.c_int * 3),
                ("start", C.c_int * 3), ("tau", C.c_double), ("U0", C.c_double), ("rho0", C.c_double),
                ("device", C.c_int), ("bcT", C.c_int * 6), ("reserved", C.c_int * 1),
                ("paraA", C.c_double), ("gBeta", C.c_double), ("Tref", C.c_double), ("Thot", C.c_double),
                ("Tcold", C.c_double), ("omegaRot", C.c_double), ("Qd", C.c_double), ("Qnu", C.c_double)]


class HaloMsg(C.Structure):
    _fields_ = [("dir", C.c_int), ("send_to", C.c_int), ("recv_from", C.c_int), ("npop", C.c_int),
                ("send_count", C.c_int), ("recv_count", C.c_int), ("pops", C.c_int * 5)]


class P2dDesc(C.Structure):
    _fields_ = [("total_nx", C.c_int), ("total_ny", C.c_int), ("nparticles", C.c_int), ("reserved", C.c_int)] + \
               [(n, C.c_double) for n in ("rho0", "rhoSolid", "viscosity", "radius0", "gravity", "thresholdWall", "stiffWall",
                                          "thresholdParticle", "stiffParticle", "Uwall", "Uframe")] + \
               [("bb_linear", C.c_int), ("moving_walls", C.c_int)]


class L2dDesc(C.Structure):
    _fields_ = [("total_nx", C.c_int), ("total_ny", C.c_int), ("variant", C.c_int), ("arith", C.c_int),
                ("reynolds", C.c_double), ("U0", C.c_double), ("rho0", C.c_double)]


class T2dDesc(C.Structure):
    _fields_ = [("total_nx", C.c_int), ("total_ny", C.c_int), ("arith", C.c_int), ("bcT", C.c_int * 4), ("variant", C.c_int),
                ("Rayleigh", C.c_double), ("Prandtl", C.c_double), ("Mach", C.c_double), ("Thot", C.c_double), ("Tcold", C.c_double),
                ("Tref", C.c_double), ("rho0", C.c_double), ("lengthUnit", C.c_double), ("Uwall", C.c_double * 8), ("cornersT", C.c_int)]


class AaDesc(C.Structure):
    _fields_ = [("n", C.c_int * 3), ("arith", C.c_int), ("collision", C.c_int), ("device", C.c_int),
                ("tau", C.c_double), ("U0", C.c_double), ("rho0", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/mglc.h one to one
SIGNATURES = {
    "mglc_version": (C.c_int, []),
    "mglc_last_error": (C.c_char_p, []),
    "mglc_strerror": (C.c_char_p, [C.c_int]),
    "mglc_device_count": (C.c_int, [_ip]),
    "mglc_dims_create": (C.c_int, [C.c_int, _ip]),
    "mglc_decompose_1d": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _ip]),
    "mglc_cart_rank": (C.c_int, [_ip, _ip, _ip]),
    "mglc_cart_coords": (C.c_int, [_ip, C.c_int, _ip]),
    "mglc_cart_neighbors": (C.c_int, [_ip, _ip, _ip, _ip]),
    "mglc_lbm_desc_init": (C.c_int, [C.POINTER(LbmDesc), _ip, _ip, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "mglc_halo_plan": (C.c_int, [C.POINTER(LbmDesc), C.POINTER(HaloMsg), _ip]),
    "mglc_relaxation_rates": (C.c_int, [C.c_double, _dp, _dp]),
    "mglc_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mglc_comm_init_rank": (C.c_int, [_vpp, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "mglc_comm_destroy": (C.c_int, [_vp]),
    "mglc_lbm_create": (C.c_int, [_vpp, C.POINTER(LbmDesc), _vp]),
    "mglc_lbm_destroy": (C.c_int, [_vp]),
    "mglc_lbm_initial": (C.c_int, [_vp]),
    "mglc_lbm_upload": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mglc_lbm_upload_fpost": (C.c_int, [_vp, _vp]),
    "mglc_lbm_download_macro": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mglc_lbm_download_f": (C.c_int, [_vp, _vp]),
    "mglc_lbm_download_fpost": (C.c_int, [_vp, _vp]),
    "mglc_collision": (C.c_int, [_vp]),
    "mglc_exchange": (C.c_int, [_vp]),
    "mglc_streaming": (C.c_int, [_vp]),
    "mglc_bounceback": (C.c_int, [_vp]),
    "mglc_macro": (C.c_int, [_vp]),
    "mglc_check": (C.c_int, [_vp, _dp]),
    "mglc_lbm_step": (C.c_int, [_vp, C.c_int]),
    "mglc_lbm_sync": (C.c_int, [_vp]),
    "mglc_lbm_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_lbm_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_lbm_kernel_time": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]),
    "mglc_lbm_set_profiling": (C.c_int, [_vp, C.c_int]),
    "mglc_lbm_set_overlap": (C.c_int, [_vp, C.c_int]),
    "mglc_lbm_get_overlap": (C.c_int, [_vp, _ip]),
    "mglc_lbm_direct_halo": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mglc_host_alloc": (C.c_int, [_vpp, C.c_size_t]),
    "mglc_host_free": (C.c_int, [_vp]),
    "mglc_lbm_device_bytes": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_lbm_get_desc": (C.c_int, [_vp, C.POINTER(LbmDesc)]),
    "mglc_group_create": (C.c_int, [_vpp, C.POINTER(LbmDesc), C.c_int, _ip]),
    "mglc_group_destroy": (C.c_int, [_vp]),
    "mglc_group_size": (C.c_int, [_vp, _ip]),
    "mglc_group_rank": (C.c_int, [_vp, C.c_int, _vpp]),
    "mglc_group_initial": (C.c_int, [_vp]),
    "mglc_group_collision": (C.c_int, [_vp]),
    "mglc_group_exchange": (C.c_int, [_vp]),
    "mglc_group_streaming": (C.c_int, [_vp]),
    "mglc_group_bounceback": (C.c_int, [_vp]),
    "mglc_group_macro": (C.c_int, [_vp]),
    "mglc_group_check": (C.c_int, [_vp, _dp]),
    "mglc_group_step": (C.c_int, [_vp, C.c_int]),
    "mglc_group_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    # thermal double-distribution path
    "mglc_thermal_desc_init": (C.c_int, [C.POINTER(LbmDesc), _ip, _ip, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "mglc_collisionT": (C.c_int, [_vp]),
    "mglc_exchange_g": (C.c_int, [_vp]),
    "mglc_streamingT": (C.c_int, [_vp]),
    "mglc_bouncebackT": (C.c_int, [_vp]),
    "mglc_macroT": (C.c_int, [_vp]),
    "mglc_check_thermal": (C.c_int, [_vp, _dp, _dp]),
    "mglc_lbm_upload_thermal": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mglc_lbm_download_thermal": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mglc_lbm_upload_gpost": (C.c_int, [_vp, _vp]),
    "mglc_lbm_download_gpost": (C.c_int, [_vp, _vp]),
    "mglc_group_collisionT": (C.c_int, [_vp]),
    "mglc_group_exchange_g": (C.c_int, [_vp]),
    "mglc_group_streamingT": (C.c_int, [_vp]),
    "mglc_group_bouncebackT": (C.c_int, [_vp]),
    "mglc_group_macroT": (C.c_int, [_vp]),
    "mglc_group_check_thermal": (C.c_int, [_vp, _dp, _dp]),
    # diagnostics and on-disk formats
    "mglc_calNuRe": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "mglc_group_calNuRe": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "mglc_lbm_download_line": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _ip, _ip]),
    "mglc_unformatted_write": (C.c_int, [C.c_char_p, C.c_int, _vpp, C.POINTER(C.c_longlong), C.c_longlong]),
    "mglc_unformatted_read": (C.c_int, [C.c_char_p, C.c_int, _vpp, C.POINTER(C.c_longlong)]),
    "mglc_output_binary_lid": (C.c_int, [C.c_char_p, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int]),
    "mglc_output_binary_thermal": (C.c_int, [C.c_char_p, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int]),
    "mglc_backup_write": (C.c_int, [C.c_char_p] + [_vp] * 6 + [C.c_int] * 3),
    "mglc_backup_read": (C.c_int, [C.c_char_p] + [_vp] * 6 + [C.c_int] * 3),
    "mglc_grid_coords": (C.c_int, [C.c_int, _vp]),
    "mglc_output_tecplot_lid": (C.c_int, [C.c_char_p] + [_vp] * 7 + [C.c_int] * 3),
    "mglc_output_tecplot_thermal": (C.c_int, [C.c_char_p] + [_vp] * 7 + [C.c_int] * 3),
    "mglc_get_velocity": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp, _vp]),
    "mglc_output_filename": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_int]),
    # Jacobi path
    "mglc_dims_create_nd": (C.c_int, [C.c_int, C.c_int, _ip]),
    "mglc_jacobi_create": (C.c_int, [_vpp, C.c_int, _ip, _ip, C.c_int, C.c_int, C.c_int, _vp]),
    "mglc_jacobi_create_local": (C.c_int, [_vpp, C.c_int, _ip, _ip, C.c_int, _ip]),
    "mglc_jacobi_destroy": (C.c_int, [_vp]),
    "mglc_jacobi_nlocal": (C.c_int, [_vp, _ip]),
    "mglc_jacobi_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "mglc_jacobi_init": (C.c_int, [_vp]),
    "mglc_jacobi_upload": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "mglc_jacobi_download": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "mglc_jacobi_exchange": (C.c_int, [_vp]),
    "mglc_jacobi_sweep": (C.c_int, [_vp]),
    "mglc_jacobi_step": (C.c_int, [_vp, C.c_int]),
    "mglc_jacobi_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_jacobi_check_diff": (C.c_int, [_vp, _dp]),
    "mglc_jacobi_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_jacobi_sync": (C.c_int, [_vp]),
    "mglc_jacobi_set_halo": (C.c_int, [_vp, C.c_int]),
    "mglc_jacobi_direct_halo": (C.c_int, [_vp, _ip]),
    # particle-laden D2Q9 path
    "mglc_p2d_desc_init": (C.c_int, [C.POINTER(P2dDesc), C.c_int]),
    "mglc_p2d_dims_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip]),
    "mglc_p2d_create": (C.c_int, [_vpp, C.POINTER(P2dDesc), _ip, C.c_int, C.c_int, C.c_int, _vp]),
    "mglc_p2d_create_local": (C.c_int, [_vpp, C.POINTER(P2dDesc), _ip, C.c_int, _ip]),
    "mglc_p2d_destroy": (C.c_int, [_vp]),
    "mglc_p2d_nlocal": (C.c_int, [_vp, _ip]),
    "mglc_p2d_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "mglc_p2d_set_particles": (C.c_int, [_vp] + [_vp] * 6),
    "mglc_p2d_get_particles": (C.c_int, [_vp] + [_vp] * 8),
    "mglc_p2d_set_forces": (C.c_int, [_vp] + [_vp] * 3),
    "mglc_p2d_upload": (C.c_int, [_vp, C.c_int] + [_vp] * 6),
    "mglc_p2d_download": (C.c_int, [_vp, C.c_int] + [_vp] * 6),
    "mglc_p2d_initial": (C.c_int, [_vp]),
    "mglc_p2d_collision": (C.c_int, [_vp]),
    "mglc_p2d_send_all_fp": (C.c_int, [_vp]),
    "mglc_p2d_streaming": (C.c_int, [_vp]),
    "mglc_p2d_bounceback": (C.c_int, [_vp]),
    "mglc_p2d_bounceback_particle": (C.c_int, [_vp, C.c_int]),
    "mglc_p2d_macro": (C.c_int, [_vp]),
    "mglc_p2d_calforce": (C.c_int, [_vp]),
    "mglc_p2d_send_all_f": (C.c_int, [_vp]),
    "mglc_p2d_update_center": (C.c_int, [_vp]),
    "mglc_p2d_check": (C.c_int, [_vp, _dp]),
    "mglc_p2d_step": (C.c_int, [_vp, C.c_int]),
    "mglc_p2d_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_p2d_set_rho_avg": (C.c_int, [_vp, C.c_double]),
    "mglc_p2d_get_rho_avg": (C.c_int, [_vp, _dp]),
    "mglc_p2d_error_flags": (C.c_int, [_vp, _ip]),
    "mglc_p2d_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_p2d_sync": (C.c_int, [_vp]),
    # 2-D D2Q9 lid-driven cavity
    "mglc_output_binary_thermal2d": (C.c_int, [C.c_char_p, _vp, _vp, _vp, C.c_int, C.c_int]),
    "mglc_backup_write_2d": (C.c_int, [C.c_char_p] + [_vp] * 5 + [C.c_int, C.c_int]),
    "mglc_backup_read_2d": (C.c_int, [C.c_char_p] + [_vp] * 5 + [C.c_int, C.c_int]),
    "mglc_halo_plan_2d": (C.c_int, [C.c_int, C.c_int, _ip, C.c_int, C.POINTER(HaloMsg), _ip]),
    "mglc_l2d_msg_table": (C.c_int, [C.c_int, C.c_int, _ip, C.c_int, C.POINTER(HaloMsg)]),
    "mglc_t2d_msg_table": (C.c_int, [C.c_int, C.c_int, _ip, C.c_int, C.POINTER(HaloMsg)]),
    "mglc_l2d_desc_init": (C.c_int, [C.POINTER(L2dDesc), C.c_int]),
    "mglc_l2d_create": (C.c_int, [_vpp, C.POINTER(L2dDesc), _ip, C.c_int, C.c_int, C.c_int, _vp]),
    "mglc_l2d_create_local": (C.c_int, [_vpp, C.POINTER(L2dDesc), _ip, C.c_int, _ip]),
    "mglc_l2d_destroy": (C.c_int, [_vp]),
    "mglc_l2d_nlocal": (C.c_int, [_vp, _ip]),
    "mglc_l2d_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "mglc_l2d_params": (C.c_int, [_vp, _dp, _dp, _dp]),
    "mglc_l2d_upload": (C.c_int, [_vp, C.c_int] + [_vp] * 5),
    "mglc_l2d_download": (C.c_int, [_vp, C.c_int] + [_vp] * 5),
    "mglc_l2d_initial": (C.c_int, [_vp]),
    "mglc_l2d_collision": (C.c_int, [_vp]),
    "mglc_l2d_exchange": (C.c_int, [_vp]),
    "mglc_l2d_streaming": (C.c_int, [_vp]),
    "mglc_l2d_bounceback": (C.c_int, [_vp]),
    "mglc_l2d_macro": (C.c_int, [_vp]),
    "mglc_l2d_check": (C.c_int, [_vp, _dp]),
    "mglc_l2d_step": (C.c_int, [_vp, C.c_int]),
    "mglc_l2d_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_l2d_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_l2d_sync": (C.c_int, [_vp]),
    "mglc_t2d_desc_init": (C.c_int, [C.POINTER(T2dDesc)]),
    "mglc_t2d_desc_init_acc": (C.c_int, [C.POINTER(T2dDesc)]),
    "mglc_t2d_desc_init_sheared_rb": (C.c_int, [C.POINTER(T2dDesc)]),
    "mglc_t2d_create": (C.c_int, [_vpp, C.POINTER(T2dDesc), _ip, C.c_int, C.c_int, C.c_int, _vp]),
    "mglc_t2d_create_local": (C.c_int, [_vpp, C.POINTER(T2dDesc), _ip, C.c_int, _ip]),
    "mglc_t2d_destroy": (C.c_int, [_vp]),
    "mglc_t2d_nlocal": (C.c_int, [_vp, _ip]),
    "mglc_t2d_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "mglc_t2d_params": (C.c_int, [_vp, _dp]),
    "mglc_t2d_upload": (C.c_int, [_vp, C.c_int] + [_vp] * 4 + [_vpp]),
    "mglc_t2d_download": (C.c_int, [_vp, C.c_int] + [_vp] * 4 + [_vpp]),
    "mglc_t2d_initial": (C.c_int, [_vp]),
    "mglc_t2d_collision": (C.c_int, [_vp]),
    "mglc_t2d_exchange_f": (C.c_int, [_vp]),
    "mglc_t2d_streaming": (C.c_int, [_vp]),
    "mglc_t2d_bounceback": (C.c_int, [_vp]),
    "mglc_t2d_collisionT": (C.c_int, [_vp]),
    "mglc_t2d_exchange_g": (C.c_int, [_vp]),
    "mglc_t2d_streamingT": (C.c_int, [_vp]),
    "mglc_t2d_bouncebackT": (C.c_int, [_vp]),
    "mglc_t2d_macro": (C.c_int, [_vp]),
    "mglc_t2d_macroT": (C.c_int, [_vp]),
    "mglc_t2d_check": (C.c_int, [_vp, _dp, _dp]),
    "mglc_t2d_nure": (C.c_int, [_vp, _dp]),
    "mglc_t2d_step": (C.c_int, [_vp, C.c_int]),
    "mglc_t2d_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_t2d_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_t2d_sync": (C.c_int, [_vp]),
    "mglc_aa_desc_init": (C.c_int, [C.POINTER(AaDesc), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "mglc_aa_create": (C.c_int, [_vpp, C.POINTER(AaDesc)]),
    "mglc_aa_destroy": (C.c_int, [_vp]),
    "mglc_aa_initial": (C.c_int, [_vp]),
    "mglc_aa_upload": (C.c_int, [_vp] + [_vp] * 5),
    "mglc_aa_step": (C.c_int, [_vp, C.c_int]),
    "mglc_aa_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_aa_check": (C.c_int, [_vp, _dp]),
    "mglc_aa_download_macro": (C.c_int, [_vp] + [_vp] * 4),
    "mglc_aa_download_f": (C.c_int, [_vp, _vp]),
    "mglc_aa_device_bytes": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_aa_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mglc_aa_sync": (C.c_int, [_vp]),
    "mglc_aa_get_block": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mglc_aa_group_create": (C.c_int, [_vpp, C.POINTER(AaDesc), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mglc_aa_group_destroy": (C.c_int, [_vp]),
    "mglc_aa_group_size": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mglc_aa_group_dims": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mglc_aa_group_rank": (C.c_int, [_vp, C.c_int, _vpp]),
    "mglc_aa_group_initial": (C.c_int, [_vp]),
    "mglc_aa_group_step": (C.c_int, [_vp, C.c_int]),
    "mglc_aa_group_step_timed": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "mglc_aa_group_check": (C.c_int, [_vp, _dp]),
    "mglc_aa_group_sync": (C.c_int, [_vp]),
    "mglc_aa_create_comm": (C.c_int, [_vpp, C.POINTER(AaDesc), _vp, C.POINTER(C.c_int)]),
}

_lib = None


def _preload_nccl():
    """libmglc.so needs libnccl.so.2.  A process that imports torch afterwards needs the NCCL torch was built against (bundled
    as nvidia/nccl/lib/libnccl.so.2, newer than the system's); the dynamic loader keeps whichever libnccl.so.2 came first, so
    the bundled one is loaded first when it exists -- libmglc.so only uses entry points every 2.x release has."""
    import sys
    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def lib():
    """Load libmglc.so (once).  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mglc_b200/csrc`).  mglc_b200 has no CPU/PyTorch fallback.")
        _preload_nccl()
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise MglcError(rc, lib().mglc_last_error().decode() or lib().mglc_strerror(rc).decode())
    return rc
