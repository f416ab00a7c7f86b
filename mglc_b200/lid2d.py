"""Host-side mirror of the reference's 2-D lid-driven cavity drivers over libmglc.so:
  variant "c" = MPI/Lid_driven_cavity/c/lid_driven_cavity.c, variant "f" = MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/,
  variant "i" = MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90 (the incompressible model: rho = 0 until the
  first macro(), u and v are the undivided momentum sums, check() is a ratio of sums of square roots).
Method names follow the reference subroutines (initial / collision / message_passing_sendrecv / streaming / bounceback / macro /
check); arrays cross the boundary as numpy arrays in the Fortran program's layout f(0:8,nx,ny), f_post(0:8,0:nx+1,0:ny+1),
rho,u,v(nx,ny), order="F"."""
import ctypes as C

import numpy as np

from . import _lib as L

VARIANTS = {"c": L.L2D_C, "f": L.L2D_F, "i": L.L2D_INCOMP, "s": L.L2D_C_SRT}


class LidDrivenCavity2D:
    def __init__(self, total=None, nprocs=1, dims=None, variant="f", Re=None, U0=None, rho0=None, strict=False, devices=None,
                 comm=None, device=0):
        lib = L.lib()
        d = L.L2dDesc()
        L.check(lib.mglc_l2d_desc_init(C.byref(d), VARIANTS[variant]))
        if total is not None:
            d.total_nx, d.total_ny = total
        if Re is not None:
            d.reynolds = Re
        if U0 is not None:
            d.U0 = U0
        if rho0 is not None:
            d.rho0 = rho0
        d.arith = L.ARITH_STRICT if strict else L.ARITH_FAST
        self.desc, self.variant, self.total = d, variant, (d.total_nx, d.total_ny)
        dz = (C.c_int * 2)(*(dims if dims else (0, 0)))
        self._h = C.c_void_p()
        if comm is not None:
            L.check(lib.mglc_l2d_create(C.byref(self._h), C.byref(d), dz, comm.nranks, comm.rank, comm.device, comm._h))
            self.nprocs = comm.nranks
        elif nprocs == 1:
            L.check(lib.mglc_l2d_create(C.byref(self._h), C.byref(d), dz, 1, 0, device, None))
            self.nprocs = 1
        else:
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_l2d_create_local(C.byref(self._h), C.byref(d), dz, nprocs, dev))
            self.nprocs = nprocs
        n = C.c_int()
        L.check(lib.mglc_l2d_nlocal(self._h, C.byref(n)))
        self.nlocal = n.value
        self.info = []
        for r in range(self.nlocal):
            dd, ln, st, co = ((C.c_int * 2)() for _ in range(4))
            nb = (C.c_int * 8)()
            L.check(lib.mglc_l2d_info(self._h, r, dd, ln, st, co, nb))
            self.dims = tuple(dd)
            self.info.append(dict(n=tuple(ln), start=tuple(st), coords=tuple(co), nbr=tuple(nb)))
        t, a, b = C.c_double(), C.c_double(), C.c_double()
        L.check(lib.mglc_l2d_params(self._h, C.byref(t), C.byref(a), C.byref(b)))
        self.tauf, self.Snu, self.Sq = t.value, a.value, b.value

    def close(self):
        if self._h:
            L.lib().mglc_l2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's subroutines ----
    def check(self):
        e = C.c_double()
        L.check(L.lib().mglc_l2d_check(self._h, C.byref(e)))
        return e.value

    def step(self, n=1):
        L.check(L.lib().mglc_l2d_step(self._h, n))

    def step_timed(self, n=1):
        ms = C.c_float()
        L.check(L.lib().mglc_l2d_step_timed(self._h, n, C.byref(ms)))
        return ms.value

    def sync(self):
        L.check(L.lib().mglc_l2d_sync(self._h))

    def launch_count(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_l2d_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- arrays ----
    def _shape(self, r, name):
        nx, ny = self.info[r]["n"]
        return {"f": (9, nx, ny), "f_post": (9, nx + 2, ny + 2)}.get(name, (nx, ny))

    _ORDER = ("f", "f_post", "rho", "u", "v")

    def upload(self, r, **arrays):
        args = []
        for name in self._ORDER:
            a = arrays.pop(name, None)
            if a is not None:
                a = np.asfortranarray(a, dtype=np.float64)
                if a.shape != self._shape(r, name):
                    raise ValueError(f"{name}: expected shape {self._shape(r, name)}, got {a.shape}")
            args.append(a)
        if arrays:
            raise TypeError(f"unknown arrays {sorted(arrays)}")
        L.check(L.lib().mglc_l2d_upload(self._h, r, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in args]))

    def download(self, r, *names):
        out = {n: np.empty(self._shape(r, n), order="F") for n in names}
        L.check(L.lib().mglc_l2d_download(self._h, r, *[out[n].ctypes.data_as(C.c_void_p) if n in out else None for n in self._ORDER]))
        return out[names[0]] if len(names) == 1 else tuple(out[n] for n in names)

    def gather(self, name):
        """Global interior array assembled from the subdomains this handle owns (NaN elsewhere)."""
        lead = (9,) if name in ("f", "f_post") else ()
        out = np.full(lead + self.total, np.nan, order="F")
        for r, inf in enumerate(self.info):
            a = self.download(r, name)
            if name == "f_post":
                a = a[:, 1:-1, 1:-1]
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            out[(slice(None),) * len(lead) + sl] = a
        return out

    def scatter(self, name, glob):
        lead = 1 if name == "f" else 0
        for r, inf in enumerate(self.info):
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            self.upload(r, **{name: glob[(slice(None),) * lead + sl]})


for _name, _sub in (("initial", "mglc_l2d_initial"), ("collision", "mglc_l2d_collision"), ("message_passing_sendrecv", "mglc_l2d_exchange"),
                    ("streaming", "mglc_l2d_streaming"), ("bounceback", "mglc_l2d_bounceback"), ("macro", "mglc_l2d_macro")):
    setattr(LidDrivenCavity2D, _name, (lambda sub: lambda self: L.check(getattr(L.lib(), sub)(self._h)))(_sub))
