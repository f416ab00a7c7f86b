"""Host-side mirror of the reference's Jacobi driver (MPI/Laplace/fortran/jacobi2d_mpi.f90, "LAP") over
libmglc.so.  Method names follow the reference subroutines: init / exchange_message / jacobi /
check_diff; arrays cross the boundary as numpy arrays in the reference layout A(0:nx+1, 0:ny+1[, 0:nz+1]),
order="F".  ndim=3 is the six-neighbour extension of BASELINE.json config 2."""
import ctypes as C

import numpy as np

from . import _lib as L


def dims_create_nd(nranks, ndim):
    d = (C.c_int * 3)()
    L.check(L.lib().mglc_dims_create_nd(nranks, ndim, d))
    return tuple(d)


class Jacobi:
    def __init__(self, total, nprocs=1, dims=None, devices=None, comm=None, device=0):
        lib = L.lib()
        self.ndim = len(total)
        if self.ndim not in (2, 3):
            raise ValueError("total must have 2 or 3 entries")
        self.total = tuple(total)
        gn = (C.c_int * 3)(*(tuple(total) + (1,))[:3])
        dz = (C.c_int * 3)(*(tuple(dims) + (1,))[:3]) if dims else (C.c_int * 3)(0, 0, 0)
        self._h = C.c_void_p()
        if comm is not None:
            L.check(lib.mglc_jacobi_create(C.byref(self._h), self.ndim, gn, dz, comm.nranks, comm.rank, comm.device, comm._h))
            self.nprocs = comm.nranks
        elif nprocs == 1:
            L.check(lib.mglc_jacobi_create(C.byref(self._h), self.ndim, gn, dz, 1, 0, device, None))
            self.nprocs = 1
        else:
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_jacobi_create_local(C.byref(self._h), self.ndim, gn, dz, nprocs, dev))
            self.nprocs = nprocs
        n = C.c_int()
        L.check(lib.mglc_jacobi_nlocal(self._h, C.byref(n)))
        self.nlocal = n.value
        self.info = []
        for r in range(self.nlocal):
            d, ln, st, co = ((C.c_int * 3)() for _ in range(4))
            nb = (C.c_int * 6)()
            L.check(lib.mglc_jacobi_info(self._h, r, d, ln, st, co, nb))
            self.dims = tuple(d)[:self.ndim]
            self.info.append(dict(n=tuple(ln)[:self.ndim], start=tuple(st)[:self.ndim], coords=tuple(co)[:self.ndim],
                                  nbr=tuple(nb)[:2 * self.ndim]))

    def close(self):
        if self._h:
            L.lib().mglc_jacobi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, r):
        return tuple(n + 2 for n in self.info[r]["n"])

    # ---- the reference's subroutines ----
    def init(self):
        L.check(L.lib().mglc_jacobi_init(self._h))

    def exchange_message(self):
        L.check(L.lib().mglc_jacobi_exchange(self._h))

    def jacobi(self):
        """jacobi(A, A_new) followed by the driver's role swap (LAP:97-103)."""
        L.check(L.lib().mglc_jacobi_sweep(self._h))

    def check_diff(self):
        e = C.c_double()
        L.check(L.lib().mglc_jacobi_check_diff(self._h, C.byref(e)))
        return e.value

    def step(self, nits=1):
        L.check(L.lib().mglc_jacobi_step(self._h, nits))

    def step_timed(self, nits=1):
        ms = C.c_float()
        L.check(L.lib().mglc_jacobi_step_timed(self._h, nits, C.byref(ms)))
        return ms.value

    def sync(self):
        L.check(L.lib().mglc_jacobi_sync(self._h))

    def set_halo(self, mode):
        """'direct' (boundary values stored into the neighbours' ghost layers by the sweep) or 'exchange' (LAP:94-103 as written)"""
        L.check(L.lib().mglc_jacobi_set_halo(self._h, {"direct": 1, "exchange": 0}[mode]))

    def direct_halo_available(self):
        a = C.c_int()
        L.check(L.lib().mglc_jacobi_direct_halo(self._h, C.byref(a)))
        return bool(a.value)

    def halo_mode(self):
        return "none (1 subdomain)" if self.nprocs == 1 else ("direct stores into the neighbours' ghost layers" if self.direct_halo_available() else "NCCL exchange, then sweep")

    def launch_count(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_jacobi_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- arrays ----
    def upload(self, r, A=None, A_new=None, f=None):
        arrs = []
        for a in (A, A_new, f):
            if a is not None:
                a = np.asfortranarray(a, dtype=np.float64)
                if a.shape != self._shape(r):
                    raise ValueError(f"expected shape {self._shape(r)}, got {a.shape}")
            arrs.append(a)
        L.check(L.lib().mglc_jacobi_upload(self._h, r, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]))

    def download(self, r, which="A"):
        out = np.empty(self._shape(r), order="F")
        p = out.ctypes.data_as(C.c_void_p)
        L.check(L.lib().mglc_jacobi_download(self._h, r, p if which == "A" else None, p if which == "A_new" else None))
        return out

    def gather(self):
        """Global interior field assembled from the subdomains this handle owns."""
        out = np.full(self.total, np.nan, order="F")
        for r, inf in enumerate(self.info):
            sl = tuple(slice(s, s + n) for s, n in zip(inf["start"], inf["n"]))
            inner = tuple(slice(1, n + 1) for n in inf["n"])
            out[sl] = self.download(r)[inner]
        return out
