"""The drivers' on-disk formats through the C ABI (mglc_b200/csrc/formats.cpp): output_binary(), backupData(),
output_Tecplot(), getVelocity() and the file names they build.  Names follow L3/output.f90 and
MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90 ("B3").  Host-only: needs no GPU."""
import ctypes as C
import os

import numpy as np

from . import _lib as L

FILE_LID_PLT, FILE_LID_BIN, FILE_LID_DAT, FILE_THERMAL_PLT, FILE_THERMAL_BIN, FILE_BACKUP = range(6)


def _f(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _path(p):
    return os.fsencode(p)


def output_filename(kind, itc):
    """'MRTcavity-'//B2//'.plt' (i9.9), 'MRTcavity-<itc>.bin', 'buoyancyCavity-<itc>.bin|.plt', 'backupFile-<itc>.bin'"""
    buf = C.create_string_buffer(128)
    L.check(L.lib().mglc_output_filename(buf, 128, kind, itc))
    return buf.value.decode()


def grid_coords(total_n):
    """xp(0:total_n+1) -- L3/initial.f90:18-31"""
    xp = np.empty(total_n + 2)
    L.check(L.lib().mglc_grid_coords(total_n, _p(xp)))
    return xp


def unformatted_write(path, records, max_subrecord=0):
    """One Fortran `write(unit) list` per array in `records` (sequential, unformatted, gfortran framing)."""
    arrs = [np.asfortranarray(r) for r in records]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    sizes = (C.c_longlong * n)(*[a.nbytes for a in arrs])
    L.check(L.lib().mglc_unformatted_write(_path(path), n, ptrs, sizes, max_subrecord))


def unformatted_read(path, shapes, dtype=np.float64):
    """Reads one record per shape (column-major), the way the reference's `read(01) (((...)))` lists do."""
    arrs = [np.empty(s, dtype=dtype, order="F") for s in shapes]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    sizes = (C.c_longlong * n)(*[a.nbytes for a in arrs])
    L.check(L.lib().mglc_unformatted_read(_path(path), n, ptrs, sizes))
    return arrs


def output_binary_lid(path, u, v, rho):
    """output_binary(), L3/output.f90:350-367: records u, v, rho"""
    u = _f(u); v = _f(v, u.shape); rho = _f(rho, u.shape)
    L.check(L.lib().mglc_output_binary_lid(_path(path), _p(u), _p(v), _p(rho), *u.shape))


def output_binary_thermal(path, u, v, w, T):
    """output_binary(), B3:1593-1618: records u, v, w, T"""
    u = _f(u); v = _f(v, u.shape); w = _f(w, u.shape); T = _f(T, u.shape)
    L.check(L.lib().mglc_output_binary_thermal(_path(path), _p(u), _p(v), _p(w), _p(T), *u.shape))


def backup_write(path, u, v, w, T, f, g):
    """backupData(), B3/seq/bouyancy3d.F90:1011-1029"""
    u = _f(u); n = u.shape
    v = _f(v, n); w = _f(w, n); T = _f(T, n); f = _f(f, (19,) + n); g = _f(g, (7,) + n)
    L.check(L.lib().mglc_backup_write(_path(path), _p(u), _p(v), _p(w), _p(T), _p(f), _p(g), *n))


def backup_read(path, n):
    """initial() with loadInitField = 1, B3/seq/bouyancy3d.F90:367-378 -> dict(u, v, w, T, f, g)"""
    n = tuple(n)
    out = {k: np.empty(n, order="F") for k in ("u", "v", "w", "T")}
    out["f"] = np.empty((19,) + n, order="F")
    out["g"] = np.empty((7,) + n, order="F")
    L.check(L.lib().mglc_backup_read(_path(path), *[_p(out[k]) for k in ("u", "v", "w", "T", "f", "g")], *n))
    return out


def output_binary_thermal2d(path, u, v, T):
    """output_binary() of the 2-D thermal driver, mpi_blocked/output.F90:192-217: records u, v, T"""
    u = _f(u); v = _f(v, u.shape); T = _f(T, u.shape)
    L.check(L.lib().mglc_output_binary_thermal2d(_path(path), _p(u), _p(v), _p(T), *u.shape))


def backup_write_2d(path, f, g, u, v, T):
    """backupData() of the 2-D thermal driver: records f, g, u, v, T as the driver stores them (f (9,nx,ny) for the MPI program,
    (nx,ny,9) for the OpenACC one) -- mpi_blocked/output.F90:381-401, seq/bouyancy2d_acc.F90:1158-1189"""
    u = _f(u); n = u.shape
    v = _f(v, n); T = _f(T, n); f = _f(f); g = _f(g)
    if f.size != 9 * u.size or g.size != 5 * u.size:
        raise ValueError("f / g do not hold 9 / 5 populations per node")
    L.check(L.lib().mglc_backup_write_2d(_path(path), _p(f), _p(g), _p(u), _p(v), _p(T), *n))


def backup_read_2d(path, n, population_last=False):
    """initial() with loadInitField = 1 -> dict(f, g, u, v, T); population_last: the OpenACC program's f(nx,ny,0:8)"""
    n = tuple(n)
    out = {k: np.empty(n, order="F") for k in ("u", "v", "T")}
    out["f"] = np.empty(n + (9,) if population_last else (9,) + n, order="F")
    out["g"] = np.empty(n + (5,) if population_last else (5,) + n, order="F")
    L.check(L.lib().mglc_backup_read_2d(_path(path), *[_p(out[k]) for k in ("f", "g", "u", "v", "T")], *n))
    return out


def output_tecplot_lid(path, u, v, w, rho):
    """output_Tecplot(), L3/output.f90:175-313 (X Y Z U V W Pressure = rho/3, all float, POINT packing)"""
    u = _f(u); n = u.shape
    v = _f(v, n); w = _f(w, n); rho = _f(rho, n)
    xp, yp, zp = (grid_coords(m) for m in n)
    L.check(L.lib().mglc_output_tecplot_lid(_path(path), _p(xp), _p(yp), _p(zp), _p(u), _p(v), _p(w), _p(rho), *n))


def output_tecplot_thermal(path, u, v, w, T):
    """output_Tecplot(), B3:1623-1773 (X Y Z U V W T)"""
    u = _f(u); n = u.shape
    v = _f(v, n); w = _f(w, n); T = _f(T, n)
    xp, yp, zp = (grid_coords(m) for m in n)
    L.check(L.lib().mglc_output_tecplot_thermal(_path(path), _p(xp), _p(yp), _p(zp), _p(u), _p(v), _p(w), _p(T), *n))


def get_velocity(u, w, U0):
    """getVelocity(), L3/output.f90:318-347 -> (u(nxHalf,nyHalf,:)/U0, zp/nz, xp/nx, w(:,nyHalf,nzHalf)/U0)"""
    u = _f(u); nx, ny, nz = u.shape
    w = _f(w, u.shape)
    xp, zp = grid_coords(nx), grid_coords(nz)
    uz, zn, xn, wx = np.empty(nz), np.empty(nz), np.empty(nx), np.empty(nx)
    L.check(L.lib().mglc_get_velocity(_p(xp), _p(zp), _p(u), _p(w), nx, ny, nz, U0, _p(uz), _p(zn), _p(xn), _p(wx)))
    return uz, zn, xn, wx
