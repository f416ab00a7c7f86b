"""Host-side mirror of the reference's lid-driven-cavity driver interface over libmglc.so.

The reference driver (MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/main.f90:85-103) calls
argument-less subroutines `collision / message_passing_sendrecv / streaming / bounceback / macro /
check` over module globals.  `LidDrivenCavity` keeps those names and meanings; the globals live on the
GPU behind an opaque handle.  Arrays cross this boundary as numpy arrays in the reference's Fortran
layout: f (19,nx,ny,nz), f_post (19,nx+2,ny+2,nz+2), rho/u/v/w (nx,ny,nz), order="F".

Three ways to run, all through the same C ABI:
  * one subdomain on one GPU                          LidDrivenCavity(total)
  * P subdomains driven by this process               LidDrivenCavity(total, nprocs=P[, devices=[...]])
  * one process per GPU (torchrun / mpirun) over NCCL LidDrivenCavity(total, comm=Communicator(...))
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _farray(a, shape, name):
    a = np.asarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return np.asfortranarray(a)


def dims_create(nranks):
    """MPI_Dims_create(nranks, 3, dims) as the reference uses it (L3/main.f90:24)."""
    d = (C.c_int * 3)()
    L.check(L.lib().mglc_dims_create(nranks, d))
    return tuple(d)


def decompose_1d(total_n, rank, nranks):
    """decompose_1d (L3/main.f90:144-155) -> (local_n, start)."""
    n, s = C.c_int(), C.c_int()
    L.check(L.lib().mglc_decompose_1d(total_n, rank, nranks, C.byref(n), C.byref(s)))
    return n.value, s.value


def cart_neighbors(dims, coords):
    """(nbr_surface {1..6}, nbr_line {7..18}) as in L3/main.f90:43-47,158-212; -1 = MPI_PROC_NULL."""
    ns, nl = (C.c_int * 6)(), (C.c_int * 12)()
    L.check(L.lib().mglc_cart_neighbors((C.c_int * 3)(*dims), (C.c_int * 3)(*coords), ns, nl))
    return {i + 1: ns[i] for i in range(6)}, {a: nl[a - 7] for a in range(7, 19)}


def make_desc(total, nranks=1, rank=0, dims=None, Re=1000.0, U0=0.1, rho0=1.0, arith="fast", device=0,
              kernel=L.KERNEL_AUTO, collision="mrt"):
    d = L.LbmDesc()
    dz = (C.c_int * 3)(*(dims if dims else (0, 0, 0)))
    L.check(L.lib().mglc_lbm_desc_init(C.byref(d), (C.c_int * 3)(*total), dz, nranks, rank, Re, U0, rho0))
    d.arith = L.ARITH_STRICT if arith == "strict" else L.ARITH_FAST
    d.device = device
    d.kernel = kernel
    if collision not in ("mrt", "bgk"):
        raise ValueError(f"collision must be 'mrt' or 'bgk', not {collision!r}")
    if collision == "bgk":          # the alternative operator of L3/collision.f90:191-198
        d.collision = L.BGK
    return d


def make_thermal_desc(total, nranks=1, rank=0, dims=None, Rayleigh=1e6, Prandtl=0.71, Mach=0.1, Ekman=0.001, bcT=None,
                      arith="fast", device=0, param_nz=None):
    """Descriptor of the buoyancy-driven cavity the way B3's module commondata computes it (B3:26-47,73-74).
    param_nz: evaluate tau, paraA, gBeta, omegaRot for a cavity of that height instead of total[2] (tests)."""
    d = L.LbmDesc()
    dz = (C.c_int * 3)(*(dims if dims else (0, 0, 0)))
    L.check(L.lib().mglc_thermal_desc_init(C.byref(d), (C.c_int * 3)(*total), dz, nranks, rank, Rayleigh, Prandtl, Mach, Ekman))
    if param_nz is not None:
        q = L.LbmDesc()
        L.check(L.lib().mglc_thermal_desc_init(C.byref(q), (C.c_int * 3)(1, 1, param_nz), (C.c_int * 3)(1, 1, 1), 1, 0,
                                               Rayleigh, Prandtl, Mach, Ekman))
        for k in ("tau", "paraA", "gBeta", "omegaRot"):
            setattr(d, k, getattr(q, k))
    d.arith = L.ARITH_STRICT if arith == "strict" else L.ARITH_FAST
    d.device = device
    if bcT is not None:
        for q in range(6):
            d.bcT[q] = bcT[q]
    return d


def halo_plan(desc):
    msgs = (L.HaloMsg * 18)()
    n = C.c_int()
    L.check(L.lib().mglc_halo_plan(C.byref(desc), msgs, C.byref(n)))
    return [dict(dir=m.dir, send_to=m.send_to, recv_from=m.recv_from, npop=m.npop, send_count=m.send_count,
                 recv_count=m.recv_count, pops=[p for p in m.pops if p >= 0]) for m in msgs[:n.value]]


def halo_plan_2d(total, dims, rank):
    """The 12 messages of the 2-D drivers for the rank at `rank` of a dims[0] x dims[1] grid (host-only, no GPU needed)"""
    msgs = (L.HaloMsg * 12)()
    n = C.c_int()
    L.check(L.lib().mglc_halo_plan_2d(total[0], total[1], (C.c_int * 2)(*dims), rank, msgs, C.byref(n)))
    return [dict(dir=m.dir, send_to=m.send_to, recv_from=m.recv_from, npop=m.npop, send_count=m.send_count,
                 recv_count=m.recv_count, pops=[p for p in m.pops if p >= 0]) for m in msgs[:n.value]]


class Communicator:
    """NCCL communicator for the one-process-per-GPU mode (replaces MPI_Cart_create, L3/main.f90:25).

    `bcast(bytes_or_None) -> bytes` must broadcast rank 0's 128-byte id to every rank (MPI_Bcast in a
    Fortran driver; torch.distributed in bench.py)."""

    def __init__(self, nranks, rank, device, bcast):
        lib = L.lib()
        buf = C.create_string_buffer(128)
        if rank == 0:
            L.check(lib.mglc_comm_unique_id(buf))
        uid = bcast(buf.raw if rank == 0 else None)
        self._h = C.c_void_p()
        L.check(lib.mglc_comm_init_rank(C.byref(self._h), uid, nranks, rank, device))
        self.nranks, self.rank, self.device = nranks, rank, device

    def close(self):
        if self._h:
            L.lib().mglc_comm_destroy(self._h)
            self._h = None


class Subdomain:
    """One rank's block: array transfers in the reference layout."""

    def __init__(self, handle):
        self._h = handle
        d = L.LbmDesc()
        L.check(L.lib().mglc_lbm_get_desc(handle, C.byref(d)))
        self.desc = d
        self.n = tuple(d.ln)
        self.start = tuple(d.start)
        self.coords = tuple(d.coords)

    def upload(self, f=None, rho=None, u=None, v=None, w=None):
        nx, ny, nz = self.n
        f = None if f is None else _farray(f, (19, nx, ny, nz), "f")
        fields = [None if a is None else _farray(a, self.n, k) for k, a in zip("rho u v w".split(), (rho, u, v, w))]
        L.check(L.lib().mglc_lbm_upload(self._h, _ptr(f), *[_ptr(a) for a in fields]))

    def upload_fpost(self, f_post):
        nx, ny, nz = self.n
        fp = _farray(f_post, (19, nx + 2, ny + 2, nz + 2), "f_post")
        L.check(L.lib().mglc_lbm_upload_fpost(self._h, _ptr(fp)))

    def download_macro(self):
        out = [np.empty(self.n, order="F") for _ in range(4)]
        L.check(L.lib().mglc_lbm_download_macro(self._h, *[_ptr(a) for a in out]))
        return dict(zip(("rho", "u", "v", "w"), out))

    def download_f(self):
        out = np.empty((19,) + self.n, order="F")
        L.check(L.lib().mglc_lbm_download_f(self._h, _ptr(out)))
        return out

    def download_fpost(self):
        nx, ny, nz = self.n
        out = np.empty((19, nx + 2, ny + 2, nz + 2), order="F")
        L.check(L.lib().mglc_lbm_download_fpost(self._h, _ptr(out)))
        return out

    # ---- thermal handles ----
    def upload_thermal(self, g=None, T=None, Fx=None, Fy=None, Fz=None):
        nx, ny, nz = self.n
        g = None if g is None else _farray(g, (7, nx, ny, nz), "g")
        fields = [None if a is None else _farray(a, self.n, k) for k, a in zip("T Fx Fy Fz".split(), (T, Fx, Fy, Fz))]
        L.check(L.lib().mglc_lbm_upload_thermal(self._h, _ptr(g), *[_ptr(a) for a in fields]))

    def download_thermal(self, with_g=True):
        out = {k: np.empty(self.n, order="F") for k in ("T", "Fx", "Fy", "Fz")}
        g = np.empty((7,) + self.n, order="F") if with_g else None
        L.check(L.lib().mglc_lbm_download_thermal(self._h, _ptr(g), *[_ptr(out[k]) for k in ("T", "Fx", "Fy", "Fz")]))
        if with_g:
            out["g"] = g
        return out

    def upload_gpost(self, g_post):
        nx, ny, nz = self.n
        gp = _farray(g_post, (7, nx + 2, ny + 2, nz + 2), "g_post")
        L.check(L.lib().mglc_lbm_upload_gpost(self._h, _ptr(gp)))

    def download_gpost(self):
        nx, ny, nz = self.n
        out = np.empty((7, nx + 2, ny + 2, nz + 2), order="F")
        L.check(L.lib().mglc_lbm_download_gpost(self._h, _ptr(out)))
        return out

    def launch_count(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_lbm_launch_count(self._h, C.byref(n)))
        return n.value

    def device_bytes(self):
        n = C.c_longlong()
        L.check(L.lib().mglc_lbm_device_bytes(self._h, C.byref(n)))
        return n.value


class LidDrivenCavity:
    """D3Q19 MRT lid-driven cavity on B200(s); method names follow L3/main.f90:85-103."""

    def __init__(self, total, nprocs=1, dims=None, Re=1000.0, U0=0.1, rho0=1.0, arith="fast", device=0,
                 devices=None, comm=None, kernel=L.KERNEL_AUTO, collision="mrt"):
        lib = L.lib()
        self.total = tuple(total)
        self.U0, self.Re, self.rho0 = U0, Re, rho0
        self._group = None
        self._single = None
        self._comm = comm
        if comm is not None:
            d = make_desc(total, comm.nranks, comm.rank, dims, Re, U0, rho0, arith, comm.device, kernel, collision)
            h = C.c_void_p()
            L.check(lib.mglc_lbm_create(C.byref(h), C.byref(d), comm._h))
            self._single = h
            self.ranks = [Subdomain(h)]
            self.nprocs = comm.nranks
        elif nprocs == 1:
            d = make_desc(total, 1, 0, dims, Re, U0, rho0, arith, device, kernel, collision)
            h = C.c_void_p()
            L.check(lib.mglc_lbm_create(C.byref(h), C.byref(d), None))
            self._single = h
            self.ranks = [Subdomain(h)]
            self.nprocs = 1
        else:
            d = make_desc(total, nprocs, 0, dims, Re, U0, rho0, arith, device, kernel, collision)
            g = C.c_void_p()
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_group_create(C.byref(g), C.byref(d), nprocs, dev))
            self._group = g
            self.ranks = []
            for r in range(nprocs):
                h = C.c_void_p()
                L.check(lib.mglc_group_rank(g, r, C.byref(h)))
                self.ranks.append(Subdomain(h))
            self.nprocs = nprocs
        self.dims = tuple(self.ranks[0].desc.dims)
        self.tauf = self.ranks[0].desc.tau

    # ---- lifetime ----
    def close(self):
        lib = L.lib()
        if self._group:
            lib.mglc_group_destroy(self._group)
            self._group = None
        elif self._single:
            lib.mglc_lbm_destroy(self._single)
            self._single = None
        self.ranks = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, single, group, *args):
        lib = L.lib()
        if self._group:
            return L.check(getattr(lib, group)(self._group, *args))
        return L.check(getattr(lib, single)(self._single, *args))

    # ---- the reference's subroutines ----
    def initial(self):
        self._call("mglc_lbm_initial", "mglc_group_initial")

    def collision(self):
        self._call("mglc_collision", "mglc_group_collision")

    def message_passing_sendrecv(self):
        self._call("mglc_exchange", "mglc_group_exchange")

    def streaming(self):
        self._call("mglc_streaming", "mglc_group_streaming")

    def bounceback(self):
        self._call("mglc_bounceback", "mglc_group_bounceback")

    def macro(self):
        self._call("mglc_macro", "mglc_group_macro")

    def check(self):
        e = C.c_double()
        self._call("mglc_check", "mglc_group_check", C.byref(e))
        return e.value

    # ---- fused fast path: n iterations of the loop body ----
    def step(self, n=1):
        self._call("mglc_lbm_step", "mglc_group_step", n)

    def step_timed(self, n=1):
        ms = C.c_float()
        self._call("mglc_lbm_step_timed", "mglc_group_step_timed", n, C.byref(ms))
        return ms.value

    def sync(self):
        for R in self.ranks:
            L.check(L.lib().mglc_lbm_sync(R._h))

    def launch_count(self):
        return sum(R.launch_count() for R in self.ranks)

    def device_bytes(self):
        return sum(R.device_bytes() for R in self.ranks)

    # ---- global-array convenience (what output() gathers, L3/output.f90:12-119) ----
    def gather(self, name):
        lead = (19,) if name in ("f",) else ()
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            blk = R.download_f() if name == "f" else R.download_macro()[name]
            out[(slice(None),) * len(lead) + sl] = blk
        return out

    def gather_macro(self):
        out = {k: np.empty(self.total, order="F") for k in ("rho", "u", "v", "w")}
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            m = R.download_macro()
            for k in out:
                out[k][sl] = m[k]
        return out

    def scatter(self, f=None, rho=None, u=None, v=None, w=None):
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            R.upload(None if f is None else f[(slice(None),) + sl],
                     *[None if a is None else a[sl] for a in (rho, u, v, w)])


    # ---- output(): gather + the reference's files (L3/output.f90:12-132) ----
    def download_line(self, field, axis, g1, g2):
        """One line of rho|u|v|w|T along `axis` through global 1-based (g1, g2), read from the device without a
        full-field download (getVelocity's profiles, L3/output.f90:334-344)."""
        fid = {"rho": 0, "u": 1, "v": 2, "w": 3, "T": 4}[field]
        out = np.full(self.total[axis], np.nan)
        for R in self.ranks:
            buf = np.empty(R.n[axis])
            first, count = C.c_int(), C.c_int()
            L.check(L.lib().mglc_lbm_download_line(R._h, fid, axis, g1, g2, _ptr(buf), C.byref(first), C.byref(count)))
            if count.value:
                out[first.value - 1:first.value - 1 + count.value] = buf[:count.value]
        return out

    def getVelocity(self):
        """u(nxHalf,nyHalf,:)/U0 and w(:,nyHalf,nzHalf)/U0 with their coordinates, L3/output.f90:318-347"""
        from . import formats as F
        nx, ny, nz = self.total
        h = [(n - 1) // 2 + 1 for n in self.total]
        uz = self.download_line("u", 2, h[0], h[1]) / self.U0
        wx = self.download_line("w", 0, h[1], h[2]) / self.U0
        return uz, F.grid_coords(nz)[1:-1] / float(nz), F.grid_coords(nx)[1:-1] / float(nx), wx

    def output(self, directory, itc, binary=True):
        """output(): MRTcavity-<itc 9 digits>.plt (+ MRTcavity-<itc>.bin); returns the paths written"""
        import os
        from . import formats as F
        m = self.gather_macro()
        paths = [os.path.join(directory, F.output_filename(F.FILE_LID_PLT, itc))]
        F.output_tecplot_lid(paths[0], m["u"], m["v"], m["w"], m["rho"])
        if binary:
            paths.append(os.path.join(directory, F.output_filename(F.FILE_LID_BIN, itc)))
            F.output_binary_lid(paths[1], m["u"], m["v"], m["rho"])
        return paths


class BuoyancyDrivenCavity(LidDrivenCavity):
    """Thermal double-distribution cavity (D3Q19 MRT flow + D3Q7 MRT temperature, Boussinesq + Coriolis) on
    B200(s); method names follow MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90:222-248."""

    def __init__(self, total, nprocs=1, dims=None, Rayleigh=1e6, Prandtl=0.71, Mach=0.1, Ekman=0.001, bcT=None,
                 arith="fast", device=0, devices=None, comm=None, param_nz=None):
        lib = L.lib()
        self.total = tuple(total)
        self._group = self._single = None
        self._comm = comm
        mk = lambda nr, r, dev: make_thermal_desc(total, nr, r, dims, Rayleigh, Prandtl, Mach, Ekman, bcT, arith, dev, param_nz)
        if comm is not None:
            d = mk(comm.nranks, comm.rank, comm.device)
            h = C.c_void_p()
            L.check(lib.mglc_lbm_create(C.byref(h), C.byref(d), comm._h))
            self._single, self.ranks, self.nprocs = h, [Subdomain(h)], comm.nranks
        elif nprocs == 1:
            d = mk(1, 0, device)
            h = C.c_void_p()
            L.check(lib.mglc_lbm_create(C.byref(h), C.byref(d), None))
            self._single, self.ranks, self.nprocs = h, [Subdomain(h)], 1
        else:
            d = mk(nprocs, 0, device)
            g = C.c_void_p()
            dev = (C.c_int * nprocs)(*devices) if devices else None
            L.check(lib.mglc_group_create(C.byref(g), C.byref(d), nprocs, dev))
            self._group, self.ranks, self.nprocs = g, [], nprocs
            for r in range(nprocs):
                h = C.c_void_p()
                L.check(lib.mglc_group_rank(g, r, C.byref(h)))
                self.ranks.append(Subdomain(h))
        self.desc = self.ranks[0].desc
        self.dims = tuple(self.desc.dims)
        self.tauf = self.desc.tau

    # f side: collision / f_message_passing_sendrecv / streaming / bounceback / macro are inherited
    def f_message_passing_sendrecv(self):
        self.message_passing_sendrecv()

    def collisionT(self):
        self._call("mglc_collisionT", "mglc_group_collisionT")

    def g_message_passing_sendrecv(self):
        self._call("mglc_exchange_g", "mglc_group_exchange_g")

    def streamingT(self):
        self._call("mglc_streamingT", "mglc_group_streamingT")

    def bouncebackT(self):
        self._call("mglc_bouncebackT", "mglc_group_bouncebackT")

    def macroT(self):
        self._call("mglc_macroT", "mglc_group_macroT")

    def check(self):
        eu, et = C.c_double(), C.c_double()
        self._call("mglc_check_thermal", "mglc_group_check_thermal", C.byref(eu), C.byref(et))
        return eu.value, et.value

    def gather(self, name):
        if name in ("rho", "u", "v", "w", "f"):
            return super().gather(name)
        lead = (7,) if name == "g" else ()
        out = np.empty(lead + self.total, order="F")
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            out[(slice(None),) * len(lead) + sl] = R.download_thermal(with_g=(name == "g"))[name]
        return out

    def gather_macro(self):
        out = super().gather_macro()
        out["T"] = self.gather("T")
        return out

    def scatter_thermal(self, g=None, T=None, Fx=None, Fy=None, Fz=None):
        for R in self.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            R.upload_thermal(None if g is None else g[(slice(None),) + sl],
                             *[None if a is None else a[sl] for a in (T, Fx, Fy, Fz)])

    # ---- diagnostics and files (B3/mpi_blocked/RaNu.F90, B3:1473-1755, seq backupData) ----
    def calNuRe(self, Prandtl=0.71):
        """(NuVolAvg, ReVolAvg) of the current fields, reduced on the device -- RaNu.F90:13-47"""
        nu, re = C.c_double(), C.c_double()
        self._call("mglc_calNuRe", "mglc_group_calNuRe", Prandtl, C.byref(nu), C.byref(re))
        return nu.value, re.value

    def output(self, directory, itc):
        """output(): buoyancyCavity-<itc>.bin and .plt, B3:1577-1578"""
        import os
        from . import formats as F
        m = self.gather_macro()
        paths = [os.path.join(directory, F.output_filename(k, itc)) for k in (F.FILE_THERMAL_BIN, F.FILE_THERMAL_PLT)]
        F.output_binary_thermal(paths[0], m["u"], m["v"], m["w"], m["T"])
        F.output_tecplot_thermal(paths[1], m["u"], m["v"], m["w"], m["T"])
        return paths

    def backupData(self, directory, itc):
        """backupFile-<itc>.bin: u, v, w, T, f, g -- B3/seq/bouyancy3d.F90:1011-1029"""
        import os
        from . import formats as F
        m = self.gather_macro()
        path = os.path.join(directory, F.output_filename(F.FILE_BACKUP, itc))
        F.backup_write(path, m["u"], m["v"], m["w"], m["T"], self.gather("f"), self.gather("g"))
        return path

    def loadInitField(self, path, rho="reference"):
        """initial() with loadInitField = 1 (B3/seq/bouyancy3d.F90:289,367-391): u,v,w,T,f,g from the backup.  The reference
        does not store rho and restarts with rho = 1 (:289) until the first macro(); rho="macro" instead recomputes it
        from f in macro()'s summation order (:739), which makes the restart continue bit for bit."""
        from . import formats as F
        b = F.backup_read(path, self.total)
        if rho == "macro":
            r = np.zeros(self.total, order="F")
            for a in range(19):
                r += b["f"][a]
        else:
            r = np.ones(self.total, order="F")
        self.scatter(f=b["f"], rho=r, u=b["u"], v=b["v"], w=b["w"])
        self.scatter_thermal(g=b["g"], T=b["T"])
        return b
