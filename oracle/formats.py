"""CPU restatement of the reference's on-disk formats (TEST INFRASTRUCTURE: only tests/ may import this).

numpy/struct restatements of
  * output_binary()   MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked/output.f90:350-367  (L3)
                      MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90:1593-1618  (B3)
  * backupData()      MPI/Buoyancy_driven_cavity/fortran/3d/seq/bouyancy3d.F90:1011-1029
  * output_Tecplot()  L3/output.f90:175-313, B3:1623-1773
  * getVelocity()     L3/output.f90:318-347
The record framing of `form="unformatted", access="sequential"` is the compiler's, not the reference's: gfortran
(the reference's Makefiles use mpif90 = gfortran) writes [int32 n][payload][int32 n] and splits records above
2 147 483 639 bytes into subrecords.  Pin: scipy.io.FortranFile, an independent reader/writer of that framing
(tests/test_formats.py); the Tecplot byte stream is pinned to the statement-by-statement list below.
"""
import struct

import numpy as np

MAX_SUBRECORD = 2147483639


def fortran_record(payload: bytes, max_subrecord=MAX_SUBRECORD) -> bytes:
    """One `write(unit) list` statement."""
    out, left, first, pos = [], len(payload), True, 0
    while True:
        n = min(left, max_subrecord)
        more = left > n
        out.append(struct.pack("<i", -n if more else n))
        out.append(payload[pos:pos + n])
        out.append(struct.pack("<i", n if first else -n))
        pos += n; left -= n; first = False
        if left <= 0:
            break
    return b"".join(out)


def _col(a):
    """(((a(i,j,k), i=1,nx), j=1,ny), k=1,nz): column-major element order"""
    return np.asarray(a, dtype="<f8").tobytes(order="F")


def output_binary_lid(u, v, rho):
    return b"".join(fortran_record(_col(a)) for a in (u, v, rho))          # L3/output.f90:362-364


def output_binary_thermal(u, v, w, T):
    return b"".join(fortran_record(_col(a)) for a in (u, v, w, T))         # B3:1611-1614


def backup_data(u, v, w, T, f, g):
    return b"".join(fortran_record(_col(a)) for a in (u, v, w, T, f, g))   # seq:1021-1026


def output_binary_thermal2d(u, v, T):
    return b"".join(fortran_record(_col(a)) for a in (u, v, T))               # B2 mpi_blocked/output.F90:211-213


def backup_data_2d(f, g, u, v, T):
    return b"".join(fortran_record(_col(a)) for a in (f, g, u, v, T))         # output.F90:397-401 ; acc:1177-1181


def grid_coords(n):
    xp = np.empty(n + 2)
    xp[0] = 0.0
    xp[n + 1] = float(n)
    for i in range(1, n + 1):
        xp[i] = float(i) - 0.5
    return xp


def _dumpstring(s):
    """dumpstring(): one default integer per character of the trimmed string, then 0 -- L3/output.f90:298-313"""
    return b"".join(struct.pack("<i", ord(c)) for c in s.rstrip(" ")) + struct.pack("<i", 0)


def output_tecplot(u, v, w, s7, name7, div3):
    """output_Tecplot(), access='stream': every `write(41)` appends the raw bytes of its item (L3/output.f90:189-293)."""
    i4 = lambda x: struct.pack("<i", x)
    r4 = lambda x: struct.pack("<f", x)
    nx, ny, nz = u.shape
    b = [b"#!TDV101", i4(1), _dumpstring("MyFirst"), i4(7)]
    for name in ("X", "Y", "Z", "U", "V", "W", name7):
        b.append(_dumpstring(name))
    b += [r4(299.0), _dumpstring("ZONE 001"), i4(-1), i4(0), i4(1), i4(0), i4(0), i4(nx), i4(ny), i4(nz), i4(0), r4(357.0)]
    b += [r4(299.0)] + [i4(1)] * 7 + [i4(0), i4(-1)]
    xp, yp, zp = grid_coords(nx), grid_coords(ny), grid_coords(nz)
    rows = np.empty((nx, ny, nz, 7), dtype="<f4")                # real(x): fp64 -> fp32, round to nearest
    rows[..., 0] = xp[1:-1, None, None]
    rows[..., 1] = yp[None, 1:-1, None]
    rows[..., 2] = zp[None, None, 1:-1]
    rows[..., 3], rows[..., 4], rows[..., 5] = u, v, w
    rows[..., 6] = (np.asarray(s7) / 3.0) if div3 else s7        # real(rho(i,j,k)/3.0d0): divide in fp64, then round
    b.append(np.transpose(rows, (2, 1, 0, 3)).tobytes(order="C"))   # do k / do j / do i / 7 items
    return b"".join(b)


def get_velocity(u, w, U0):
    nx, ny, nz = u.shape
    nxHalf, nyHalf, nzHalf = (nx - 1) // 2 + 1, (ny - 1) // 2 + 1, (nz - 1) // 2 + 1
    xp, zp = grid_coords(nx), grid_coords(nz)
    uz = np.array([u[nxHalf - 1, nyHalf - 1, k - 1] / U0 for k in range(1, nz + 1)])
    zn = np.array([zp[k] / float(nz) for k in range(1, nz + 1)])
    xn = np.array([xp[i] / float(nx) for i in range(1, nx + 1)])
    wx = np.array([w[i - 1, nyHalf - 1, nzHalf - 1] / U0 for i in range(1, nx + 1)])
    return uz, zn, xn, wx
