"""Host-side mirror of the reference's 2-D thermal driver (MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/) over libmglc.so.
Method names follow the reference subroutines (initial / collision / message_passing_f / streaming / bounceback / collisionT /
message_passing_g / streamingT / bouncebackT / macro / macroT / check); arrays cross the boundary as numpy arrays in the Fortran
program's layout f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1), g(0:4,nx,ny), g_post(0:4,0:nx+1,0:ny+1), rho,u,v,T,Fx,Fy(nx,ny), order="F"."""
import ctypes as C

import numpy as np

from . import _lib as L

SIDE_HEATED = (L.BCT_CONST_COLD, L.BCT_CONST_HOT, L.BCT_ADIABATIC, L.BCT_ADIABATIC)      # +x, -x, +y, -y   macros.F90:24-27
RAYLEIGH_BENARD = (L.BCT_ADIABATIC, L.BCT_ADIABATIC, L.BCT_CONST_COLD, L.BCT_CONST_HOT)  # macros.F90:17-20
RB_PERIODIC = (L.BCT_PERIODIC, L.BCT_PERIODIC, L.BCT_CONST_COLD, L.BCT_CONST_HOT)        # seq/bouyancy2d_acc.F90:13-22 (the OpenACC program)
PARAM_NAMES = ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu", "lengthUnit")


class BuoyancyDrivenCavity2D:
    LEAD = {"f": 9, "f_post": 9, "g": 5, "g_post": 5}
    _LATTICES = ("f", "f_post", "g", "g_post")
    _FIELDS = ("rho", "u", "v", "T", "Fx", "Fy")

    def __init__(self, total=None, nprocs=1, dims=None, bcT=None, strict=False, devices=None, comm=None, device=0, variant="mpi",
                 lengthUnit=None, Uwall=None, cornersT=None, shearReynolds=None, **params):
        """variant "mpi": mpi_blocked/ (201 x 201, Ra 1e7, side-heated); "acc": the OpenACC program seq/bouyancy2d_acc.F90
        (513 x 257, Ra 1e5, Rayleigh-Benard plates, periodic vertical walls, lengthUnit = nx); "sheared_rb": seq/R_B_2d.F90
        (201 x 201, Ra 1e7, Pr 5.3, Rayleigh-Benard plates, walls moving at shearReynolds = 100, its corner cells in bouncebackT())
        -- each with its shipped defaults.  Uwall = (TopLeft, TopRight, BottomLeft, BottomRight, LeftTop, LeftBottom, RightTop,
        RightBottom) sets the wall velocities directly; shearReynolds = R sets them as R_B_2d.F90:118-120 does (U0 = R*viscosity/ny)."""
        lib = L.lib()
        d = L.T2dDesc()
        init = {"acc": lib.mglc_t2d_desc_init_acc, "sheared_rb": lib.mglc_t2d_desc_init_sheared_rb}.get(variant, lib.mglc_t2d_desc_init)
        L.check(init(C.byref(d)))
        if variant == "sheared_rb" and shearReynolds is None and Uwall is None:
            shearReynolds = 100.0                      # R_B_2d.F90:79 (recomputed below for the lattice and parameters of this run)
        if variant == "acc" and total is not None and lengthUnit is None:
            lengthUnit = float(total[0])              # acc:57: lengthUnit = dble(nx)
        if lengthUnit is not None:
            d.lengthUnit = lengthUnit
        if total is not None:
            d.total_nx, d.total_ny = total
        if bcT is not None:
            d.bcT[:] = list(bcT)
        for k, v in params.items():
            if k not in ("Rayleigh", "Prandtl", "Mach", "Thot", "Tcold", "Tref", "rho0"):
                raise TypeError(f"unknown parameter {k}")
            setattr(d, k, v)
        if shearReynolds is not None:                 # R_B_2d.F90:58,96-97,118-120, the same products in the same order
            import math
            lu = d.lengthUnit if d.lengthUnit > 0.0 else float(d.total_ny)
            tauf = 0.5 + d.Mach * lu * math.sqrt(3.0 * d.Prandtl / d.Rayleigh)
            U0 = shearReynolds * ((tauf - 0.5) / 3.0) / float(d.total_ny)
            Uwall = (U0, -U0, -U0, U0, U0, U0, U0, U0)
        if Uwall is not None:
            d.Uwall[:] = [float(x) for x in Uwall]
        if cornersT is not None:
            d.cornersT = int(bool(cornersT))
        d.arith = L.ARITH_STRICT if strict else L.ARITH_FAST
        self.desc, self.total, self.bcT, self.variant = d, (d.total_nx, d.total_ny), tuple(d.bcT), variant
        self.Uwall, self.cornersT = tuple(d.Uwall), bool(d.cornersT)
        dz = (C.c_int * 2)(*(dims if dims else (0, 0)))
        self._h = C.c_void_p()
        if comm is not None:
            L.check(lib.mglc_t2d_create(C.byref(self._h), C.byref(d), dz, comm.nranks, comm.rank, comm.device, comm._h))
            self.nprocs = comm.nranks
        elif nprocs == 1:
            L.check(lib.mglc_t2d_create(C.byref(self._h), C.byref(d), dz, 1, 0, device, None))
            self.nprocs = 1
        else:
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_t2d_create_local(C.byref(self._h), C.byref(d), dz, nprocs, dev))
            self.nprocs = nprocs
        n = C.c_int()
        L.check(lib.mglc_t2d_nlocal(self._h, C.byref(n)))
        self.nlocal = n.value
        self.info = []
        for r in range(self.nlocal):
            dd, ln, st, co = ((C.c_int * 2)() for _ in range(4))
            nb = (C.c_int * 8)()
            L.check(lib.mglc_t2d_info(self._h, r, dd, ln, st, co, nb))
            self.dims = tuple(dd)
            self.info.append(dict(n=tuple(ln), start=tuple(st), coords=tuple(co), nbr=tuple(nb)))
        out = (C.c_double * 10)()
        L.check(lib.mglc_t2d_params(self._h, out))
        self.params = dict(zip(PARAM_NAMES, out))

    def close(self):
        if self._h:
            L.lib().mglc_t2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's subroutines ----
    def check(self):
        eu, et = C.c_double(), C.c_double()
        L.check(L.lib().mglc_t2d_check(self._h, C.byref(eu), C.byref(et)))
        return eu.value, et.value

    def calNuRe(self):
        """(angular momentum / N, NuVolAvg, ReVolAvg) -- NuRe.F90:27-78"""
        out = (C.c_double * 3)()
        L.check(L.lib().mglc_t2d_nure(self._h, out))
        return tuple(out)

    def step(self, n=1):
        L.check(L.lib().mglc_t2d_step(self._h, n))

    def step_timed(self, n=1):
        ms = C.c_float()
        L.check(L.lib().mglc_t2d_step_timed(self._h, n, C.byref(ms)))
        return ms.value

    def sync(self):
        L.check(L.lib().mglc_t2d_sync(self._h))

    def launch_count(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_t2d_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- arrays ----
    def _shape(self, r, name):
        nx, ny = self.info[r]["n"]
        return {"f": (9, nx, ny), "f_post": (9, nx + 2, ny + 2), "g": (5, nx, ny), "g_post": (5, nx + 2, ny + 2)}.get(name, (nx, ny))

    def upload(self, r, **arrays):
        keep = {}
        for name, a in arrays.items():
            if name not in self._LATTICES + self._FIELDS:
                raise TypeError(f"unknown array {name}")
            a = np.asfortranarray(a, dtype=np.float64)
            if a.shape != self._shape(r, name):
                raise ValueError(f"{name}: expected shape {self._shape(r, name)}, got {a.shape}")
            keep[name] = a
        ptr = lambda n: keep[n].ctypes.data_as(C.c_void_p) if n in keep else None
        fields = (C.c_void_p * 6)(*[ptr(n) for n in self._FIELDS])
        L.check(L.lib().mglc_t2d_upload(self._h, r, *[ptr(n) for n in self._LATTICES], fields))

    def download(self, r, *names):
        out = {n: np.empty(self._shape(r, n), order="F") for n in names}
        ptr = lambda n: out[n].ctypes.data_as(C.c_void_p) if n in out else None
        fields = (C.c_void_p * 6)(*[ptr(n) for n in self._FIELDS])
        L.check(L.lib().mglc_t2d_download(self._h, r, *[ptr(n) for n in self._LATTICES], fields))
        return out[names[0]] if len(names) == 1 else tuple(out[n] for n in names)

    def gather(self, name):
        """Global interior array assembled from the subdomains this handle owns (NaN elsewhere)."""
        lead = (self.LEAD[name],) if name in self.LEAD else ()
        out = np.full(lead + self.total, np.nan, order="F")
        for r, inf in enumerate(self.info):
            a = self.download(r, name)
            if name in ("f_post", "g_post"):
                a = a[:, 1:-1, 1:-1]
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            out[(slice(None),) * len(lead) + sl] = a
        return out

    def scatter(self, name, glob):
        lead = 1 if name in self.LEAD else 0
        for r, inf in enumerate(self.info):
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            self.upload(r, **{name: glob[(slice(None),) * lead + sl]})


for _name, _sub in (("initial", "mglc_t2d_initial"), ("collision", "mglc_t2d_collision"), ("message_passing_f", "mglc_t2d_exchange_f"),
                    ("streaming", "mglc_t2d_streaming"), ("bounceback", "mglc_t2d_bounceback"), ("collisionT", "mglc_t2d_collisionT"),
                    ("message_passing_g", "mglc_t2d_exchange_g"), ("streamingT", "mglc_t2d_streamingT"),
                    ("bouncebackT", "mglc_t2d_bouncebackT"), ("macro", "mglc_t2d_macro"), ("macroT", "mglc_t2d_macroT")):
    setattr(BuoyancyDrivenCavity2D, _name, (lambda sub: lambda self: L.check(getattr(L.lib(), sub)(self._h)))(_sub))
