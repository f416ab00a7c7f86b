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
    return lib


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

    def __init__(self, total, nprocs=1, dims=None, Re=1000.0, U0=0.1, rho0=1.0, native=False):
        self._lib = load(native) if native else lib()
        d = (C.c_int * 3)(*(dims if dims else (0, 0, 0)))
        self._h = self._lib.orc_world_create(total[0], total[1], total[2], nprocs, d, Re, U0, rho0)
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
