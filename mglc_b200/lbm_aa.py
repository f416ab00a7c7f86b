"""Host-side mirror of the reference's D3Q19 lid-driven cavity driver (L3/main.f90) on ONE lattice (AA-pattern storage,
mglc_aa_* in include/mglc.h): same loop body, same results as LidDrivenCavity, half the lattice memory.  One subdomain, or
(nranks > 1) the blocks of mpi_starts/decompose_1d (L3/main.f90:33-63) inside this process on devices that can address each
other (mglc_aa_group_*): the global arrays are scattered to / gathered from the blocks here.
Arrays cross the boundary in the Fortran program's layout f(0:18,nx,ny,nz), rho,u,v,w(nx,ny,nz), order="F"."""
import ctypes as C

import numpy as np

from . import _lib as L


class LidDrivenCavityAA:
    """LidDrivenCavityAA(total): one subdomain.  LidDrivenCavityAA(total, nranks=P, dims=None, devices=None): P blocks in this
    process.  LidDrivenCavityAA(total, comm=Communicator(...), dims=None): this process holds one block of the global lattice
    `total` (one process per GPU); every method is then a collective call, arrays are the block's (self.n, self.start)."""

    def __new__(cls, total, *args, nranks=1, **kw):
        if cls is LidDrivenCavityAA and nranks > 1:
            return super().__new__(LidDrivenCavityAAGroup)
        return super().__new__(cls)

    def __init__(self, total, Re=1000.0, U0=0.1, rho0=1.0, arith="fast", collision="mrt", device=0, nranks=1, comm=None, dims=None):
        lib = L.lib()
        d = L.AaDesc()
        L.check(lib.mglc_aa_desc_init(C.byref(d), *total, Re, U0, rho0))
        d.arith = {"fast": L.ARITH_FAST, "strict": L.ARITH_STRICT}[arith]
        d.collision = {"mrt": L.MRT_LID, "bgk": L.BGK}[collision]
        d.device = device
        self.desc, self.total, self.tauf = d, tuple(total), d.tau
        self._h = C.c_void_p()
        self.global_total, self.start, self._comm = tuple(total), (0, 0, 0), comm
        if comm is None:
            L.check(lib.mglc_aa_create(C.byref(self._h), C.byref(d)))
        else:
            cd = (C.c_int * 3)(*dims) if dims else None
            L.check(lib.mglc_aa_create_comm(C.byref(self._h), C.byref(d), comm._h, cd))
            ln, st = (C.c_int * 3)(), (C.c_int * 3)()
            L.check(lib.mglc_aa_get_block(self._h, ln, st))
            self.total, self.start = tuple(ln), tuple(st)       # the arrays of this rank are its block's
        self.n = self.total

    def close(self):
        if self._h:
            L.lib().mglc_aa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initial(self):
        L.check(L.lib().mglc_aa_initial(self._h))

    def step(self, n=1):
        L.check(L.lib().mglc_aa_step(self._h, n))

    def step_timed(self, n=1):
        ms = C.c_float()
        L.check(L.lib().mglc_aa_step_timed(self._h, n, C.byref(ms)))
        return ms.value

    def check(self):
        e = C.c_double()
        L.check(L.lib().mglc_aa_check(self._h, C.byref(e)))
        return e.value

    def sync(self):
        L.check(L.lib().mglc_aa_sync(self._h))

    def launch_count(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_aa_launch_count(self._h, C.byref(n)))
        return n.value

    def device_bytes(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_aa_device_bytes(self._h, C.byref(n)))
        return n.value

    def upload(self, f=None, rho=None, u=None, v=None, w=None):
        keep = []
        for name, a in (("f", f), ("rho", rho), ("u", u), ("v", v), ("w", w)):
            if a is not None:
                a = np.asfortranarray(a, dtype=np.float64)
                want = ((19,) if name == "f" else ()) + self.total
                if a.shape != want:
                    raise ValueError(f"{name}: expected shape {want}, got {a.shape}")
            keep.append(a)
        L.check(L.lib().mglc_aa_upload(self._h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in keep]))

    def download_macro(self):
        out = {k: np.empty(self.total, order="F") for k in ("rho", "u", "v", "w")}
        L.check(L.lib().mglc_aa_download_macro(self._h, *[out[k].ctypes.data_as(C.c_void_p) for k in ("rho", "u", "v", "w")]))
        return out

    def download_f(self):
        f = np.empty((19,) + self.total, order="F")
        L.check(L.lib().mglc_aa_download_f(self._h, f.ctypes.data_as(C.c_void_p)))
        return f


class _Block:
    """per-block view of a group member (download / upload in the block's own shape)"""

    def __init__(self, h):
        self._h = h
        ln, st = (C.c_int * 3)(), (C.c_int * 3)()
        L.check(L.lib().mglc_aa_get_block(h, ln, st))
        self.n, self.start, self.total = tuple(ln), tuple(st), tuple(ln)

    upload = LidDrivenCavityAA.upload
    download_macro = LidDrivenCavityAA.download_macro
    download_f = LidDrivenCavityAA.download_f
    launch_count = LidDrivenCavityAA.launch_count
    device_bytes = LidDrivenCavityAA.device_bytes

    @property
    def slices(self):
        return tuple(slice(s, s + n) for s, n in zip(self.start, self.n))


class LidDrivenCavityAAGroup(LidDrivenCavityAA):
    """LidDrivenCavityAA(total, nranks=P, dims=None, devices=None): P blocks in this process, total = the GLOBAL lattice"""

    def __init__(self, total, Re=1000.0, U0=0.1, rho0=1.0, arith="fast", collision="mrt", device=0, nranks=2, dims=None, devices=None):
        lib = L.lib()
        d = L.AaDesc()
        L.check(lib.mglc_aa_desc_init(C.byref(d), *total, Re, U0, rho0))
        d.arith = {"fast": L.ARITH_FAST, "strict": L.ARITH_STRICT}[arith]
        d.collision = {"mrt": L.MRT_LID, "bgk": L.BGK}[collision]
        d.device = device
        self.desc, self.total, self.tauf = d, tuple(total), d.tau
        self._h = None
        self._g = C.c_void_p()
        cd = (C.c_int * 3)(*dims) if dims else None
        dev = (C.c_int * nranks)(*devices) if devices is not None else None
        L.check(lib.mglc_aa_group_create(C.byref(self._g), C.byref(d), nranks, cd, dev))
        got = (C.c_int * 3)()
        L.check(lib.mglc_aa_group_dims(self._g, got))
        self.dims = tuple(got)
        self.blocks = []
        for r in range(nranks):
            h = C.c_void_p()
            L.check(lib.mglc_aa_group_rank(self._g, r, C.byref(h)))
            self.blocks.append(_Block(h))

    def close(self):
        if self._g:
            L.lib().mglc_aa_group_destroy(self._g)
            self._g = None

    def initial(self):
        L.check(L.lib().mglc_aa_group_initial(self._g))

    def step(self, n=1):
        L.check(L.lib().mglc_aa_group_step(self._g, n))

    def step_timed(self, n=1):
        ms = C.c_float()
        L.check(L.lib().mglc_aa_group_step_timed(self._g, n, C.byref(ms)))
        return ms.value

    def check(self):
        e = C.c_double()
        L.check(L.lib().mglc_aa_group_check(self._g, C.byref(e)))
        return e.value

    def sync(self):
        L.check(L.lib().mglc_aa_group_sync(self._g))

    def launch_count(self):
        return sum(b.launch_count() for b in self.blocks)

    def device_bytes(self):
        return sum(b.device_bytes() for b in self.blocks)

    def upload(self, f=None, rho=None, u=None, v=None, w=None):
        self.sync()
        for b in self.blocks:
            part = [None if a is None else np.asfortranarray(np.asarray(a)[((slice(None),) if i == 0 else ()) + b.slices])
                    for i, a in enumerate((f, rho, u, v, w))]
            b.upload(*part)

    def download_macro(self):
        self.sync()
        out = {k: np.empty(self.total, order="F") for k in ("rho", "u", "v", "w")}
        for b in self.blocks:
            m = b.download_macro()
            for k in out:
                out[k][b.slices] = m[k]
        return out

    def download_f(self):
        self.sync()
        f = np.empty((19,) + self.total, order="F")
        for b in self.blocks:
            f[(slice(None),) + b.slices] = b.download_f()
        return f
