"""Host-side mirror of the reference's particle-laden channel driver (MPI/Micro_particles/fortran/case4/
mpi_particle/main.F90:35-73, "P4") over libmglc.so.  Method names follow the reference subroutines.  Arrays cross
the boundary as numpy arrays in the reference layout, order="F": f (9,nx+6,ny+6), f_post (9,nx+4,ny+4),
obst (nx+2,ny+2) int32, rho/u/v (nx,ny).  Particle positions are an input (the reference seeds them with a
compiler-specific random_number)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ParticleChannel:
    def __init__(self, x, y, radius=None, nprocs=1, dims=None, devices=None, comm=None, device=0, **params):
        lib = L.lib()
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        self.N = len(x)
        d = L.P2dDesc()
        L.check(lib.mglc_p2d_desc_init(C.byref(d), self.N))
        for k, v in params.items():
            setattr(d, k, v)
        self.desc = d
        self.total = (d.total_nx, d.total_ny)
        dz = (C.c_int * 2)(*(dims if dims else (0, 0)))
        self._h = C.c_void_p()
        if comm is not None:
            L.check(lib.mglc_p2d_create(C.byref(self._h), C.byref(d), dz, comm.nranks, comm.rank, comm.device, comm._h))
            self.nprocs = comm.nranks
        elif nprocs == 1:
            L.check(lib.mglc_p2d_create(C.byref(self._h), C.byref(d), dz, 1, 0, device, None))
            self.nprocs = 1
        else:
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_p2d_create_local(C.byref(self._h), C.byref(d), dz, nprocs, dev))
            self.nprocs = nprocs
        n = C.c_int()
        L.check(lib.mglc_p2d_nlocal(self._h, C.byref(n)))
        self.info = []
        for r in range(n.value):
            dd, ln, st, co = ((C.c_int * 2)() for _ in range(4))
            nb = (C.c_int * 8)()
            L.check(lib.mglc_p2d_info(self._h, r, dd, ln, st, co, nb))
            self.dims = tuple(dd)
            self.info.append(dict(n=tuple(ln), start=tuple(st), coords=tuple(co), nbr=tuple(nb)))
        rad = np.full(self.N, d.radius0) if radius is None else np.ascontiguousarray(radius, dtype=np.float64)
        self.set_particles(x=x, y=y, radius=rad)

    def close(self):
        if self._h:
            L.lib().mglc_p2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- particle state ----
    def set_particles(self, x=None, y=None, U=None, V=None, omega=None, radius=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, U, V, omega, radius)]
        L.check(L.lib().mglc_p2d_set_particles(self._h, *[_p(a) for a in arrs]))

    def particles(self):
        names = ("xCenter", "yCenter", "Uc", "Vc", "rationalOmega", "wallTotalForceX", "wallTotalForceY", "totalTorque")
        out = {k: np.empty(self.N) for k in names}
        L.check(L.lib().mglc_p2d_get_particles(self._h, *[_p(out[k]) for k in names]))
        return out

    def set_forces(self, Fx, Fy, torque):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Fx, Fy, torque)]
        L.check(L.lib().mglc_p2d_set_forces(self._h, *[_p(a) for a in arrs]))

    # ---- the reference's subroutines ----
    def _call(self, name, *args):
        L.check(getattr(L.lib(), name)(self._h, *args))

    def initial(self):
        self._call("mglc_p2d_initial")

    def collision(self):
        self._call("mglc_p2d_collision")

    def send_all_fp(self):
        self._call("mglc_p2d_send_all_fp")

    def streaming(self):
        self._call("mglc_p2d_streaming")

    def bounceback(self):
        self._call("mglc_p2d_bounceback")

    def bounceback_particle(self, recompute_rho_avg=True):
        self._call("mglc_p2d_bounceback_particle", 1 if recompute_rho_avg else 0)

    def macro(self):
        self._call("mglc_p2d_macro")

    def calForce(self):
        self._call("mglc_p2d_calforce")

    def send_all_f(self):
        self._call("mglc_p2d_send_all_f")

    def updateCenter(self):
        self._call("mglc_p2d_update_center")

    def check(self):
        e = C.c_double()
        self._call("mglc_p2d_check", C.byref(e))
        return e.value

    def step(self, n=1):
        self._call("mglc_p2d_step", n)

    def step_timed(self, n=1):
        ms = C.c_float()
        self._call("mglc_p2d_step_timed", n, C.byref(ms))
        return ms.value

    def sync(self):
        self._call("mglc_p2d_sync")

    def set_rho_avg(self, v):
        self._call("mglc_p2d_set_rho_avg", float(v))

    def rho_avg(self):
        e = C.c_double()
        self._call("mglc_p2d_get_rho_avg", C.byref(e))
        return e.value

    def error_flags(self):
        f = C.c_int()
        L.lib().mglc_p2d_error_flags(self._h, C.byref(f))
        return f.value

    def launch_count(self):
        n = C.c_longlong()
        self._call("mglc_p2d_launch_count", C.byref(n))
        return n.value

    # ---- arrays ----
    def shapes(self, r):
        nx, ny = self.info[r]["n"]
        return {"f": (9, nx + 6, ny + 6), "f_post": (9, nx + 4, ny + 4), "rho": (nx, ny), "u": (nx, ny), "v": (nx, ny),
                "obst": (nx + 2, ny + 2)}

    def upload(self, r, **arrays):
        sh = self.shapes(r)
        args = []
        for k in ("f", "f_post", "rho", "u", "v", "obst"):
            a = arrays.get(k)
            if a is not None:
                a = np.asfortranarray(a, dtype=np.int32 if k == "obst" else np.float64)
                if a.shape != sh[k]:
                    raise ValueError(f"{k}: expected {sh[k]}, got {a.shape}")
            args.append(a)
        L.check(L.lib().mglc_p2d_upload(self._h, r, *[_p(a) for a in args]))

    def download(self, r, names=("f", "f_post", "rho", "u", "v", "obst")):
        sh = self.shapes(r)
        out = {k: np.empty(sh[k], order="F", dtype=np.int32 if k == "obst" else np.float64) for k in names}
        L.check(L.lib().mglc_p2d_download(self._h, r, *[_p(out.get(k)) for k in ("f", "f_post", "rho", "u", "v", "obst")]))
        return out

    def gather(self, name):
        lead = (9,) if name in ("f", "f_post") else ()
        out = np.empty(lead + self.total, order="F", dtype=np.int32 if name == "obst" else np.float64)
        rim = {"f": 3, "f_post": 2, "obst": 1}.get(name, 0)
        for r, inf in enumerate(self.info):
            nx, ny = inf["n"]
            sl = (slice(inf["start"][0], inf["start"][0] + nx), slice(inf["start"][1], inf["start"][1] + ny))
            a = self.download(r, (name,))[name]
            out[(slice(None),) * len(lead) + sl] = a[(slice(None),) * len(lead) + (slice(rim, rim + nx), slice(rim, rim + ny))]
        return out
