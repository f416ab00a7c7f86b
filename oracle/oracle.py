"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/lid3d.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Arrays are exposed as numpy views in the reference's Fortran layout (order="F"):
f (19,nx,ny,nz), f_post (19,nx+2,ny+2,nz+2), rho/u/v/w (nx,ny,nz).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(native=False):
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", _HERE, target])
    return os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")


def load(native=False):
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, s)) for s in os.listdir(_HERE) if s.endswith(".c"))
    if not os.path.exists(path) or os.path.getmtime(path) < src_mtime:
        path = build(native)
    lib = C.CDLL(path)
    lib.orc_world_create.restype = C.c_void_p
    lib.orc_world_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double]
    lib.orc_world_destroy.argtypes = [C.c_void_p]
    lib.orc_rank_ptr.restype = _dp
    lib.orc_rank_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.orc_rank_info.argtypes = [C.c_void_p, C.c_int, _ip]
    lib.orc_world_info.argtypes = [C.c_void_p, _ip, _dp, _ip]
    for name in ("orc_initial", "orc_collision", "orc_exchange", "orc_streaming", "orc_bounceback", "orc_macro"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = None
    lib.orc_check.argtypes = [C.c_void_p]
    lib.orc_check.restype = C.c_double
    lib.orc_step.argtypes = [C.c_void_p, C.c_int]
    lib.orc_step.restype = None
    lib.orc_dims_create.argtypes = [C.c_int, _ip]
    lib.orc_decompose_1d.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip]
    lib.orc_moments.argtypes = [_dp, _dp]
    lib.orc_inverse.argtypes = [_dp, _dp]
    lib.orc_meq.argtypes = [C.c_double] * 4 + [_dp]
    lib.orc_collide_cell.argtypes = [_dp] + [C.c_double] * 6 + [_dp]
    lib.orc_collide_cell_bgk.argtypes = [_dp] + [C.c_double] * 5 + [_dp]
    lib.orc_world_set_bgk.argtypes = [C.c_void_p, C.c_int]
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_set_threads.restype = C.c_int
    return lib


def set_threads(n=0, native=False):
    """Run the oracle's OpenMP loops on n threads (0 = every core this process may use) and return the count the OpenMP
    runtime reports afterwards -- the number bench.py prints as `cores`.  Overrides an inherited OMP_NUM_THREADS=1."""
    if n <= 0:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    return int((load(native) if native else lib()).orc_set_threads(n))


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


EX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0])
EY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
EZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1])
OPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)


class Rank:
    """numpy views (no copies) of one emulated MPI rank's arrays."""

    def __init__(self, world, r):
        L = world._lib
        info = (C.c_int * 27)()
        L.orc_rank_info(world._h, r, info)
        self.n = tuple(info[0:3])
        self.coords = tuple(info[3:6])
        self.start = tuple(info[6:9])
        self.nbr_surface = {s: info[8 + s] for s in range(1, 7)}
        self.nbr_line = {a: info[8 + a] for a in range(7, 19)}
        nx, ny, nz = self.n

        def view(which, shape):
            p = L.orc_rank_ptr(world._h, r, which)
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape, order="F")

        self.f = view(0, (19, nx, ny, nz))
        self.f_post = view(1, (19, nx + 2, ny + 2, nz + 2))
        self.rho, self.u, self.v, self.w = (view(q, (nx, ny, nz)) for q in (2, 3, 4, 5))
        self.up, self.vp, self.wp = (view(q, (nx, ny, nz)) for q in (6, 7, 8))


class LidWorld:
    """All P emulated ranks of the lid-driven cavity (L3/main.f90) in one process."""

    def __init__(self, total, nprocs=1, dims=None, Re=1000.0, U0=0.1, rho0=1.0, native=False, collision="mrt"):
        self._lib = load(native) if native else lib()
        d = (C.c_int * 3)(*(dims if dims else (0, 0, 0)))
        self._h = self._lib.orc_world_create(total[0], total[1], total[2], nprocs, d, Re, U0, rho0)
        if collision not in ("mrt", "bgk"):
            raise ValueError(collision)
        self._lib.orc_world_set_bgk(self._h, int(collision == "bgk"))   # L3/collision.f90:191-198
        self.total = tuple(total)
        self.nprocs = nprocs
        self.U0, self.Re, self.rho0 = U0, Re, rho0
        self.ranks = [Rank(self, r) for r in range(nprocs)]
        dd = (C.c_int * 3)()
        pp = (C.c_double * 4)()
        it = C.c_int()
        self._lib.orc_world_info(self._h, dd, pp, C.byref(it))
        self.dims = tuple(dd)
        self.tauf, self.Snu, self.Sq = pp[0], pp[1], pp[2]

    def close(self):
        if self._h:
            self.ranks = []
            self._lib.orc_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initial(self):
        self._lib.orc_initial(self._h)

    def collision(self):
        self._lib.orc_collision(self._h)

    def message_passing_sendrecv(self):
        self._lib.orc_exchange(self._h)

    def streaming(self):
        self._lib.orc_streaming(self._h)

    def bounceback(self):
        self._lib.orc_bounceback(self._h)

    def macro(self):
        self._lib.orc_macro(self._h)

    def check(self):
        return self._lib.orc_check(self._h)

    def step(self, n=1):
        self._lib.orc_step(self._h, n)

    def gather(self, name):
        """Assemble a global (nx,ny,nz) field (or (19,...) for f) from the ranks, like output() does."""
        lead = (19,) if name == "f" else ()
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            out[(slice(None),) * len(lead) + sl] = getattr(R, name)
        return out

    def scatter(self, name, glob):
        lead = 1 if name == "f" else 0
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            getattr(R, name)[...] = glob[(slice(None),) * lead + sl]


def feq(rho, u, v, w):
    """Second-order equilibrium of L3/initial.f90:63-73, vectorised (rho,u,v,w broadcastable)."""
    rho, u, v, w = (np.asarray(a, dtype=np.float64) for a in (rho, u, v, w))
    us2 = u * u + v * v + w * w
    out = np.empty((19,) + np.broadcast(rho, u).shape)
    for a in range(19):
        un = u * EX[a] + v * EY[a] + w * EZ[a]
        out[a] = rho * W[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2)
    return out


# ---------------------------------------------------------------------------------------------------------
# Jacobi (oracle/jacobi.c)
def _jac_lib():
    L = lib()
    if not getattr(L, "_jac_ready", False):
        L.jac_world_create.restype = C.c_void_p
        L.jac_world_create.argtypes = [C.c_int] * 5 + [_ip]
        L.jac_world_destroy.argtypes = [C.c_void_p]
        L.jac_ptr.restype = _dp
        L.jac_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.jac_rank_info.argtypes = [C.c_void_p, C.c_int, _ip]
        L.jac_world_dims.argtypes = [C.c_void_p, _ip]
        L.jac_dims_create.argtypes = [C.c_int, C.c_int, _ip]
        for name in ("jac_init", "jac_exchange", "jac_sweep"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.jac_step.argtypes = [C.c_void_p, C.c_int]
        L.jac_step.restype = None
        L.jac_check_diff.argtypes = [C.c_void_p]
        L.jac_check_diff.restype = C.c_double
        L._jac_ready = True
    return L


class JacobiWorld:
    """All P emulated ranks of the Jacobi driver (LAP) in one process; ndim = len(total)."""

    def __init__(self, total, nprocs=1, dims=None):
        self._lib = _jac_lib()
        self.ndim = len(total)
        self.total = tuple(total)
        t = tuple(total) + (1,) * (3 - self.ndim)
        d = (C.c_int * 3)(*((tuple(dims) + (1,))[:3] if dims else (0, 0, 0)))
        self._h = self._lib.jac_world_create(self.ndim, t[0], t[1], t[2], nprocs, d)
        self.nprocs = nprocs
        dd = (C.c_int * 3)()
        self._lib.jac_world_dims(self._h, dd)
        self.dims = tuple(dd)[:self.ndim]
        self.info = []
        for r in range(nprocs):
            o = (C.c_int * 15)()
            self._lib.jac_rank_info(self._h, r, o)
            self.info.append(dict(n=tuple(o[0:3])[:self.ndim], coords=tuple(o[3:6])[:self.ndim],
                                  start=tuple(o[6:9])[:self.ndim], nbr=tuple(o[9:15])[:2 * self.ndim]))

    def close(self):
        if self._h:
            self._lib.jac_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, r, which="A"):
        """numpy view (no copy) of rank r's A / A_new / f / A_p; re-fetch after sweeps (the roles swap)."""
        shape = tuple(n + 2 for n in self.info[r]["n"])
        p = self._lib.jac_ptr(self._h, r, {"A": 0, "A_new": 1, "f": 2, "A_p": 3}[which])
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape, order="F")

    def init(self):
        self._lib.jac_init(self._h)

    def exchange_message(self):
        self._lib.jac_exchange(self._h)

    def jacobi(self):
        self._lib.jac_sweep(self._h)

    def step(self, nits=1):
        self._lib.jac_step(self._h, nits)

    def check_diff(self):
        return self._lib.jac_check_diff(self._h)

    def gather(self):
        out = np.full(self.total, np.nan, order="F")
        for r, inf in enumerate(self.info):
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            inner = tuple(slice(1, n + 1) for n in inf["n"])
            out[sl] = self.array(r)[inner]
        return out


# ---------------------------------------------------------------------------------------------------------
# Thermal double-distribution cavity (oracle/thermal3d.c)
class ThParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Rayleigh", "Prandtl", "Mach", "Ekman", "Thot", "Tcold", "Tref", "tauf",
                                          "viscosity", "diffusivity", "omegaRatating", "paraA", "gBeta1", "gBeta",
                                          "Snu", "Sq", "Qd", "Qnu")]


TH_ADIABATIC, TH_CONST_HOT, TH_CONST_COLD = 0, 1, 2
TH_FIELDS = {"f": 0, "f_post": 1, "g": 2, "g_post": 3, "rho": 4, "u": 5, "v": 6, "w": 7, "T": 8, "Fx": 9, "Fy": 10,
             "Fz": 11, "up": 12, "vp": 13, "wp": 14, "Tp": 15}


def _th_lib():
    L = lib()
    if not getattr(L, "_th_ready", False):
        L.th_world_create.restype = C.c_void_p
        L.th_world_create.argtypes = [C.c_int] * 4 + [_ip, _ip] + [C.c_double] * 4
        L.th_world_destroy.argtypes = [C.c_void_p]
        L.th_rank_ptr.restype = _dp
        L.th_rank_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.th_rank_info.argtypes = [C.c_void_p, C.c_int, _ip]
        L.th_world_info.argtypes = [C.c_void_p, _ip, C.POINTER(ThParams), _ip]
        L.th_make_params.argtypes = [C.c_int] + [C.c_double] * 7 + [C.POINTER(ThParams)]
        for name in ("th_initial", "th_collision", "th_exchange_f", "th_streaming", "th_bounceback", "th_collisionT",
                     "th_exchange_g", "th_streamingT", "th_bouncebackT", "th_macro", "th_macroT"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.th_step.argtypes = [C.c_void_p, C.c_int]
        L.th_step.restype = None
        L.th_check.argtypes = [C.c_void_p, _dp, _dp]
        L.th_feq_cell.argtypes = [C.c_double] * 4 + [_dp]
        L.th_geq_cell.argtypes = [C.c_double] * 5 + [_dp]
        L.th_collide_cell.argtypes = [_dp] + [C.c_double] * 5 + [C.POINTER(ThParams), _dp, _dp]
        L.th_collideT_cell.argtypes = [_dp] + [C.c_double] * 4 + [C.POINTER(ThParams), _dp]
        L.th_macro_cell.argtypes = [_dp] + [C.c_double] * 3 + [_dp]
        L._th_ready = True
    return L


def th_params(total_nz=51, Rayleigh=1e6, Prandtl=0.71, Mach=0.1, Ekman=0.001, Thot=1.0, Tcold=0.0, Tref=0.0):
    p = ThParams()
    _th_lib().th_make_params(total_nz, Rayleigh, Prandtl, Mach, Ekman, Thot, Tcold, Tref, C.byref(p))
    return p


class ThermalRank:
    def __init__(self, world, r):
        L = world._lib
        info = (C.c_int * 27)()
        L.th_rank_info(world._h, r, info)
        self.n = tuple(info[0:3])
        self.coords = tuple(info[3:6])
        self.start = tuple(info[6:9])
        nx, ny, nz = self.n
        shapes = {"f": (19, nx, ny, nz), "f_post": (19, nx + 2, ny + 2, nz + 2), "g": (7, nx, ny, nz),
                  "g_post": (7, nx + 2, ny + 2, nz + 2)}
        for name, which in TH_FIELDS.items():
            shape = shapes.get(name, (nx, ny, nz))
            p = L.th_rank_ptr(world._h, r, which)
            setattr(self, name, np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape, order="F"))


class ThermalWorld:
    """All P emulated ranks of the buoyancy-driven cavity (B3) in one process."""

    def __init__(self, total, nprocs=1, dims=None, bcT=None, Rayleigh=1e6, Prandtl=0.71, Mach=0.1, Ekman=0.001):
        self._lib = _th_lib()
        d = (C.c_int * 3)(*(dims if dims else (0, 0, 0)))
        bc = (C.c_int * 6)(*bcT) if bcT else None
        self._h = self._lib.th_world_create(total[0], total[1], total[2], nprocs, d, bc, Rayleigh, Prandtl, Mach, Ekman)
        self.total, self.nprocs = tuple(total), nprocs
        self.ranks = [ThermalRank(self, r) for r in range(nprocs)]
        dd, bb = (C.c_int * 3)(), (C.c_int * 6)()
        self.p = ThParams()
        self._lib.th_world_info(self._h, dd, C.byref(self.p), bb)
        self.dims, self.bcT = tuple(dd), tuple(bb)

    def close(self):
        if self._h:
            self.ranks = []
            self._lib.th_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        eu, et = C.c_double(), C.c_double()
        self._lib.th_check(self._h, C.byref(eu), C.byref(et))
        return eu.value, et.value

    def calNuRe(self, Prandtl=0.71):
        nu, re = C.c_double(), C.c_double()
        self._lib.th_calNuRe.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        self._lib.th_calNuRe(self._h, Prandtl, C.byref(nu), C.byref(re))
        return nu.value, re.value

    def step(self, n=1):
        self._lib.th_step(self._h, n)

    def gather(self, name):
        lead = {"f": (19,), "g": (7,)}.get(name, ())
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            out[(slice(None),) * len(lead) + sl] = getattr(R, name)
        return out

    def scatter(self, name, glob):
        lead = 1 if name in ("f", "g") else 0
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            getattr(R, name)[...] = glob[(slice(None),) * lead + sl]


for _name, _sub in (("initial", "th_initial"), ("collision", "th_collision"), ("f_message_passing_sendrecv", "th_exchange_f"),
                    ("streaming", "th_streaming"), ("bounceback", "th_bounceback"), ("collisionT", "th_collisionT"),
                    ("g_message_passing_sendrecv", "th_exchange_g"), ("streamingT", "th_streamingT"),
                    ("bouncebackT", "th_bouncebackT"), ("macro", "th_macro"), ("macroT", "th_macroT")):
    setattr(ThermalWorld, _name, (lambda sub: lambda self: getattr(self._lib, sub)(self._h))(_sub))


# ---------------------------------------------------------------------------------------------------------
# Particle-laden D2Q9 channel (oracle/particles2d.c)
class P2Params(C.Structure):
    _fields_ = [("total_nx", C.c_int), ("total_ny", C.c_int), ("N", C.c_int)] + \
               [(n, C.c_double) for n in ("rho0", "rhoSolid", "viscosity", "tauf", "Snu", "Sq", "gravity", "thresholdWall",
                                          "stiffWall", "thresholdParticle", "stiffParticle", "radius0", "Pi", "Uwall", "Uframe")] + \
               [("bb_linear", C.c_int), ("moving_walls", C.c_int)]


P2_FIELDS = {"f": 0, "f_post": 1, "rho": 2, "u": 3, "v": 4, "up": 5, "vp": 6}
P2_PARTICLE = {"xCenter": 0, "yCenter": 1, "Uc": 2, "Vc": 3, "rationalOmega": 4, "radius": 5, "wallTotalForceX": 6,
               "wallTotalForceY": 7, "totalTorque": 8, "xCenterOld": 9, "yCenterOld": 10, "UcOld": 11, "VcOld": 12,
               "rationalOmegaOld": 13}
EX9 = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
EY9 = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])
OPP9 = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])


def _p2_lib():
    L = lib()
    if not getattr(L, "_p2_ready", False):
        pp = C.POINTER(P2Params)
        L.p2_default_params.argtypes = [pp]
        L.p2_dims_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip]
        L.p2_world_create.restype = C.c_void_p
        L.p2_world_create.argtypes = [pp, C.c_int, _ip]
        L.p2_world_destroy.argtypes = [C.c_void_p]
        L.p2_rank_ptr.restype = _dp
        L.p2_rank_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.p2_rank_obst.restype = _ip
        L.p2_rank_obst.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.p2_rank_info.argtypes = [C.c_void_p, C.c_int, _ip]
        L.p2_particle_ptr.restype = _dp
        L.p2_particle_ptr.argtypes = [C.c_void_p, C.c_int]
        L.p2_world_info.argtypes = [C.c_void_p, _ip, _dp, _ip]
        L.p2_get_params.argtypes = [C.c_void_p, pp]
        for name in ("p2_initial", "p2_collision", "p2_send_all_fp", "p2_send_all_f", "p2_streaming", "p2_bounceback",
                     "p2_bounceback_particle", "p2_macro", "p2_calForce", "p2_updateCenter"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.p2_check.argtypes = [C.c_void_p]
        L.p2_check.restype = C.c_double
        L.p2_step.argtypes = [C.c_void_p, C.c_int]
        L.p2_step.restype = None
        L.p2_collide_cell.argtypes = [_dp] + [C.c_double] * 5 + [_dp]
        L.p2_macro_cell.argtypes = [_dp, _dp]
        L.p2_calQ.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
        L.p2_particle_forces.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.p2_particle_advance.argtypes = [pp] + [C.c_double] * 9 + [_dp]
        L.p2_set_rhoAvg.argtypes = [C.c_void_p, C.c_double]
        L.p2_bb_link_r.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.p2_force_link_r.argtypes = [C.c_void_p] + [C.c_int] * 5 + [_dp]
        L.p2_refill_cell_r.argtypes = [C.c_void_p] + [C.c_int] * 4
        L._p2_ready = True
    return L


def p2_default_params(**over):
    p = P2Params()
    _p2_lib().p2_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


class ParticleRank:
    def __init__(self, world, r):
        L = world._lib
        info = (C.c_int * 14)()
        L.p2_rank_info(world._h, r, info)
        self.n = (info[0], info[1])
        self.coords = (info[2], info[3])
        self.start = (info[4], info[5])
        self.nbr = dict(zip(("left", "right", "bottom", "top", "tl", "tr", "bl", "br"), info[6:14]))
        nx, ny = self.n
        shapes = {"f": (9, nx + 6, ny + 6), "f_post": (9, nx + 4, ny + 4)}
        for name, which in P2_FIELDS.items():
            shape = shapes.get(name, (nx, ny))
            p = L.p2_rank_ptr(world._h, r, which)
            setattr(self, name, np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape, order="F"))
        for name, which in (("obst", 0), ("obstNew", 1)):
            p = L.p2_rank_obst(world._h, r, which)
            setattr(self, name, np.ctypeslib.as_array(p, shape=((nx + 2) * (ny + 2),)).reshape((nx + 2, ny + 2), order="F"))


class ParticleWorld:
    """All P emulated ranks of the particle driver (P4/main.F90) in one process.  Particle positions are an
    input (the reference draws them from a compiler-specific random_number)."""

    def __init__(self, x, y, radius=None, nprocs=1, dims=None, **params):
        self._lib = _p2_lib()
        x = np.asarray(x, dtype=np.float64)
        self.p = p2_default_params(N=len(x), **params)
        d = (C.c_int * 2)(*(dims if dims else (0, 0)))
        self._h = self._lib.p2_world_create(C.byref(self.p), nprocs, d)
        self.nprocs, self.N = nprocs, len(x)
        self.total = (self.p.total_nx, self.p.total_ny)
        for name, which in P2_PARTICLE.items():
            ptr = self._lib.p2_particle_ptr(self._h, which)
            setattr(self, name, np.ctypeslib.as_array(ptr, shape=(self.N,)))
        self.xCenter[:] = x
        self.yCenter[:] = y
        self.radius[:] = self.p.radius0 if radius is None else radius
        self.ranks = [ParticleRank(self, r) for r in range(nprocs)]
        dd, sc, ie = (C.c_int * 2)(), (C.c_double * 2)(), (C.c_int * 2)()
        self._lib.p2_world_info(self._h, dd, sc, ie)
        self.dims = tuple(dd)

    def close(self):
        if self._h:
            self.ranks = []
            self._lib.p2_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        dd, sc, ie = (C.c_int * 2)(), (C.c_double * 2)(), (C.c_int * 2)()
        self._lib.p2_world_info(self._h, dd, sc, ie)
        return dict(rhoAvg=sc[0], errorU=sc[1], itc=ie[0], error_flag=ie[1])

    def check(self):
        return self._lib.p2_check(self._h)

    def step(self, n=1):
        self._lib.p2_step(self._h, n)

    def gather(self, name):
        out = np.empty(self.total, order="F", dtype=np.int32 if name.startswith("obst") else np.float64) if name not in ("f", "f_post") \
            else np.empty((9,) + self.total, order="F")
        for R in self.ranks:
            nx, ny = R.n
            sl = (slice(R.start[0], R.start[0] + nx), slice(R.start[1], R.start[1] + ny))
            if name == "f":
                out[(slice(None),) + sl] = R.f[:, 3:nx + 3, 3:ny + 3]
            elif name == "f_post":
                out[(slice(None),) + sl] = R.f_post[:, 2:nx + 2, 2:ny + 2]
            elif name.startswith("obst"):
                out[sl] = getattr(R, name)[1:nx + 1, 1:ny + 1]
            else:
                out[sl] = getattr(R, name)
        return out


for _name, _sub in (("initial", "p2_initial"), ("collision", "p2_collision"), ("send_all_fp", "p2_send_all_fp"),
                    ("send_all_f", "p2_send_all_f"), ("streaming", "p2_streaming"), ("bounceback", "p2_bounceback"),
                    ("bounceback_particle", "p2_bounceback_particle"), ("macro", "p2_macro"), ("calForce", "p2_calForce"),
                    ("updateCenter", "p2_updateCenter")):
    setattr(ParticleWorld, _name, (lambda sub: lambda self: getattr(self._lib, sub)(self._h))(_sub))


# ---------------------------------------------------------------------------------------------------------
# 2-D D2Q9 lid-driven cavity (oracle/lid2d.c): variant "c" = MPI/Lid_driven_cavity/c/lid_driven_cavity.c,
# variant "f" = MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked,
# variant "i" = MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90 (incompressible equilibrium),
# variant "s" = the C program with its model switch set to SRT (c:13-14, c:160-176: BGK)
L2_FIELDS = {"f": 0, "f_post": 1, "rho": 2, "u": 3, "v": 4, "up": 5, "vp": 6}
L2_VARIANTS = {"c": 0, "f": 1, "i": 2, "s": 3}


def _l2_lib():
    L = lib()
    if not getattr(L, "_l2_ready", False):
        L.l2_world_create.restype = C.c_void_p
        L.l2_world_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double]
        L.l2_world_destroy.argtypes = [C.c_void_p]
        L.l2_world_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), _dp]
        L.l2_rank_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.l2_rank_ptr.restype = _dp
        L.l2_rank_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for fn in ("l2_initial", "l2_collision", "l2_exchange", "l2_streaming", "l2_bounceback", "l2_macro"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.l2_check.restype = C.c_double
        L.l2_check.argtypes = [C.c_void_p]
        L.l2_step.argtypes = [C.c_void_p, C.c_int]
        L.l2_collide_cell.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        L.l2_dims_create.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L._l2_ready = True
    return L


class Lid2DRank:
    def __init__(self, world, r):
        L = world._lib
        info = (C.c_int * 14)()
        L.l2_rank_info(world._h, r, info)
        self.n, self.coords, self.start = tuple(info[0:2]), tuple(info[2:4]), tuple(info[4:6])
        self.nbr, self.cnr = tuple(info[6:10]), tuple(info[10:14])
        nx, ny = self.n
        shapes = {"f": (9, nx, ny), "f_post": (9, nx + 2, ny + 2)}
        for name, which in L2_FIELDS.items():
            shape = shapes.get(name, (nx, ny))
            p = L.l2_rank_ptr(world._h, r, which)
            setattr(self, name, np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape, order="F"))


class Lid2DWorld:
    """All P emulated ranks of the 2-D lid-driven cavity in one process."""

    def __init__(self, total, nprocs=1, dims=None, variant="f", Re=1000.0, U0=0.1, rho0=1.0):
        self._lib = _l2_lib()
        d = (C.c_int * 2)(*(dims if dims else (0, 0)))
        self._h = self._lib.l2_world_create(total[0], total[1], nprocs, d, L2_VARIANTS[variant], Re, U0, rho0)
        self.total, self.nprocs, self.variant = tuple(total), nprocs, variant
        self.ranks = [Lid2DRank(self, r) for r in range(nprocs)]
        dd, par = (C.c_int * 2)(), (C.c_double * 3)()
        self._lib.l2_world_info(self._h, dd, par)
        self.dims, (self.tauf, self.Snu, self.Sq) = tuple(dd), tuple(par)

    def close(self):
        if self._h:
            self.ranks = []
            self._lib.l2_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        return self._lib.l2_check(self._h)

    def step(self, n=1):
        self._lib.l2_step(self._h, n)

    def gather(self, name):
        lead = (9,) if name == "f" else ()
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            out[(slice(None),) * len(lead) + sl] = getattr(R, name)
        return out

    def scatter(self, name, glob):
        lead = 1 if name == "f" else 0
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            getattr(R, name)[...] = glob[(slice(None),) * lead + sl]


for _name, _sub in (("initial", "l2_initial"), ("collision", "l2_collision"), ("message_passing_sendrecv", "l2_exchange"),
                    ("streaming", "l2_streaming"), ("bounceback", "l2_bounceback"), ("macro", "l2_macro")):
    setattr(Lid2DWorld, _name, (lambda sub: lambda self: getattr(self._lib, sub)(self._h))(_sub))


def l2_collide_cell(variant, f, rho, u, v, Snu, Sq):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.empty(9)
    _l2_lib().l2_collide_cell(L2_VARIANTS[variant], f.ctypes.data_as(_dp), rho, u, v, Snu, Sq, out.ctypes.data_as(_dp))
    return out


class RefLid2D:
    """The reference's own compiled C program (oracle/_ref/liblid2d_ref.so, built from
    /root/reference/MPI/Lid_driven_cavity/c/lid_driven_cavity.c by `make -C oracle ref`): its functions and global arrays."""
    NX = NY = 200

    def __init__(self, path=None):
        import os
        path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liblid2d_ref.so")
        self.lib = C.CDLL(path)            # RTLD_LOCAL: its globals are named rho, u, v, f ...
        for fn in ("initial", "collision", "streaming", "boundary", "macro", "output_binary"):
            getattr(self.lib, fn).restype = C.c_int
        self.lib.check.restype = C.c_double
        self.lib.check.argtypes = [C.c_int]
        self.lib.output_tecplot.argtypes = [C.c_int]
        n = self.NX * self.NY

        def arr(name, shape):
            a = (C.c_double * int(np.prod(shape))).in_dll(self.lib, name)
            return np.ctypeslib.as_array(a).reshape(shape)            # C order: [NX][NY]([9])
        self.f, self.f_post = arr("f", (self.NX, self.NY, 9)), arr("f_post", (self.NX, self.NY, 9))
        self.rho, self.u, self.v = (arr(k, (self.NX, self.NY)) for k in ("rho", "u", "v"))
        self.up, self.vp = arr("up", (self.NX, self.NY)), arr("vp", (self.NX, self.NY))

    def scalar(self, name):
        return C.c_double.in_dll(self.lib, name).value

    def step(self, n=1):                  # the body of main()'s while loop, c:57-63
        for _ in range(n):
            self.lib.collision(); self.lib.streaming(); self.lib.boundary(); self.lib.macro()

    # the C arrays seen in the oracle's layout: f(0:8, nx, ny), rho(nx, ny)
    def f_F(self, post=False):
        return np.asfortranarray(np.transpose(self.f_post if post else self.f, (2, 0, 1)))

    def field_F(self, name):
        return np.asfortranarray(getattr(self, name))


# ---------------------------------------------------------------------------------------------------------------------
# 2-D thermal D2Q9 + D2Q5 (oracle/thermal2d.c): B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/
class T2Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Rayleigh", "Prandtl", "Mach", "Thot", "Tcold", "Tref", "rho0", "lengthUnit", "tauf",
                                          "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")]


T2_FIELDS = {"f": 0, "f_post": 1, "g": 2, "g_post": 3, "rho": 4, "u": 5, "v": 6, "T": 7, "up": 8, "vp": 9, "Tp": 10, "Fx": 11, "Fy": 12}
T2_ADIABATIC, T2_CONST_HOT, T2_CONST_COLD, T2_PERIODIC = 0, 1, 2, 3
T2_MPI, T2_ACC = 0, 1
# the OpenACC program's shipped set (seq/bouyancy2d_acc.F90:13-22): Rayleigh-Benard cell, vertical walls periodic for f and g
T2_RB_PERIODIC = (T2_PERIODIC, T2_PERIODIC, T2_CONST_COLD, T2_CONST_HOT)
T2_SIDE_HEATED = (T2_CONST_COLD, T2_CONST_HOT, T2_ADIABATIC, T2_ADIABATIC)      # +x, -x, +y, -y   macros.F90:24-27
T2_RAYLEIGH_BENARD = (T2_ADIABATIC, T2_ADIABATIC, T2_CONST_COLD, T2_CONST_HOT)  # macros.F90:17-20


def _t2_lib():
    L = lib()
    if not getattr(L, "_t2_ready", False):
        L.t2_world_create.restype = C.c_void_p
        L.t2_world_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _dp, _ip]
        L.t2_world_destroy.argtypes = [C.c_void_p]
        L.t2_world_info.argtypes = [C.c_void_p, _ip, C.POINTER(T2Params), _ip]
        L.t2_rank_info.argtypes = [C.c_void_p, C.c_int, _ip]
        L.t2_rank_ptr.restype = _dp
        L.t2_rank_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for fn in ("t2_initial", "t2_collision", "t2_exchange_f", "t2_streaming", "t2_bounceback", "t2_collisionT", "t2_exchange_g",
                   "t2_streamingT", "t2_bouncebackT", "t2_macro", "t2_macroT"):
            getattr(L, fn).argtypes = [C.c_void_p]
            getattr(L, fn).restype = None
        L.t2_check.argtypes = [C.c_void_p, _dp, _dp]
        L.t2_nure_sums.argtypes = [C.c_void_p, _dp]
        L.t2_step.argtypes = [C.c_void_p, C.c_int]
        L.t2_derive_params.argtypes = [C.POINTER(T2Params), C.c_int]
        L.t2_collide_cell.argtypes = [C.POINTER(T2Params), _dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.t2_collideT_cell.argtypes = [C.POINTER(T2Params), _dp, C.c_double, C.c_double, C.c_double, _dp]
        L.t2_collide_cell_v.argtypes = [C.c_int, C.POINTER(T2Params), _dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.t2_world_set_variant.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L._t2_ready = True
    return L


def t2_params(total_ny=201, Rayleigh=1e7, Prandtl=0.71, Mach=0.1, Thot=1.0, Tcold=0.0, Tref=0.0, rho0=1.0, lengthUnit=0.0):
    """module.F90:29-33,67-81 evaluated by the oracle (lengthUnit = 0: dble(total_ny); the OpenACC program uses dble(nx), acc:57)"""
    p = T2Params(Rayleigh=Rayleigh, Prandtl=Prandtl, Mach=Mach, Thot=Thot, Tcold=Tcold, Tref=Tref, rho0=rho0, lengthUnit=lengthUnit)
    _t2_lib().t2_derive_params(C.byref(p), total_ny)
    return p


class Thermal2DRank:
    def __init__(self, world, r):
        L = world._lib
        info = (C.c_int * 14)()
        L.t2_rank_info(world._h, r, info)
        self.n, self.coords, self.start = tuple(info[0:2]), tuple(info[2:4]), tuple(info[4:6])
        self.nbr, self.cnr = tuple(info[6:10]), tuple(info[10:14])
        nx, ny = self.n
        shapes = {"f": (9, nx, ny), "f_post": (9, nx + 2, ny + 2), "g": (5, nx, ny), "g_post": (5, nx + 2, ny + 2)}
        for name, which in T2_FIELDS.items():
            shape = shapes.get(name, (nx, ny))
            p = L.t2_rank_ptr(world._h, r, which)
            setattr(self, name, np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape, order="F"))


class Thermal2DWorld:
    """All P emulated ranks of the 2-D thermal driver in one process."""
    LEAD = {"f": 9, "f_post": 9, "g": 5, "g_post": 5}

    def __init__(self, total=(201, 201), nprocs=1, dims=None, bcT=None, variant="mpi", lengthUnit=0.0, Uwall=None, cornersT=False, **params):
        """Uwall = (TopLeft, TopRight, BottomLeft, BottomRight, LeftTop, LeftBottom, RightTop, RightBottom), cornersT: the moving
        walls and the corner rule of the sheared Rayleigh-Benard programs (seq/R_B_2d.F90:118-120, :1086-1106)"""
        self._lib = _t2_lib()
        d = (C.c_int * 2)(*(dims if dims else (0, 0)))
        pv = dict(Rayleigh=1e7, Prandtl=0.71, Mach=0.1, Thot=1.0, Tcold=0.0, Tref=0.0, rho0=1.0)
        pv.update(params)
        par = (C.c_double * 7)(*[pv[k] for k in ("Rayleigh", "Prandtl", "Mach", "Thot", "Tcold", "Tref", "rho0")])
        bc = (C.c_int * 4)(*bcT) if bcT is not None else None
        self._h = self._lib.t2_world_create(total[0], total[1], nprocs, d, par, bc)
        self.variant = variant
        if variant != "mpi" or lengthUnit:
            self._lib.t2_world_set_variant(self._h, {"mpi": T2_MPI, "acc": T2_ACC}[variant], float(lengthUnit))
        self.Uwall, self.cornersT = tuple(float(x) for x in (Uwall if Uwall is not None else [0.0] * 8)), bool(cornersT)
        if Uwall is not None or cornersT:
            self._lib.t2_world_set_walls(self._h, (C.c_double * 8)(*(Uwall if Uwall is not None else [0.0] * 8)), int(bool(cornersT)))
        self.total, self.nprocs = tuple(total), nprocs
        self.ranks = [Thermal2DRank(self, r) for r in range(nprocs)]
        dd, bb = (C.c_int * 2)(), (C.c_int * 4)()
        self.params = T2Params()
        self._lib.t2_world_info(self._h, dd, C.byref(self.params), bb)
        self.dims, self.bcT = tuple(dd), tuple(bb)
        if T2_PERIODIC in self.bcT:
            assert self.bcT[0] == self.bcT[1] == T2_PERIODIC and self.dims[0] == 1, "periodic vertical walls need bcT[0] = bcT[1] and dims[0] = 1"

    def close(self):
        if self._h:
            self.ranks = []
            self._lib.t2_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        eu, et = C.c_double(), C.c_double()
        self._lib.t2_check(self._h, C.byref(eu), C.byref(et))
        return eu.value, et.value

    def nure_sums(self):
        out = (C.c_double * 3)()
        self._lib.t2_nure_sums(self._h, out)
        return tuple(out)

    def step(self, n=1):
        self._lib.t2_step(self._h, n)

    def gather(self, name):
        lead = (self.LEAD[name],) if name in self.LEAD else ()
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            a = getattr(R, name)
            if name in ("f_post", "g_post"):
                a = a[:, 1:-1, 1:-1]
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            out[(slice(None),) * len(lead) + sl] = a
        return out

    def scatter(self, name, glob):
        lead = 1 if name in self.LEAD else 0
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            getattr(R, name)[...] = glob[(slice(None),) * lead + sl]


for _name, _sub in (("initial", "t2_initial"), ("collision", "t2_collision"), ("message_passing_f", "t2_exchange_f"),
                    ("streaming", "t2_streaming"), ("bounceback", "t2_bounceback"), ("collisionT", "t2_collisionT"),
                    ("message_passing_g", "t2_exchange_g"), ("streamingT", "t2_streamingT"), ("bouncebackT", "t2_bouncebackT"),
                    ("macro", "t2_macro"), ("macroT", "t2_macroT")):
    setattr(Thermal2DWorld, _name, (lambda sub: lambda self: getattr(self._lib, sub)(self._h))(_sub))


def t2_collide_cell(p, f, rho, u, v, T, variant="mpi"):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out, F2 = np.empty(9), np.empty(2)
    _t2_lib().t2_collide_cell_v({"mpi": T2_MPI, "acc": T2_ACC}[variant], C.byref(p), f.ctypes.data_as(_dp), rho, u, v, T,
                                out.ctypes.data_as(_dp), F2.ctypes.data_as(_dp))
    return out, F2


def t2_collideT_cell(p, g, u, v, T):
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.empty(5)
    _t2_lib().t2_collideT_cell(C.byref(p), g.ctypes.data_as(_dp), u, v, T, out.ctypes.data_as(_dp))
    return out
