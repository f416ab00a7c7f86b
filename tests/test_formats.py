"""On-disk formats of the drivers' output() / backupData(): libmglc.so's host-side writers against
the numpy restatement (oracle/formats.py), both pinned to scipy.io.FortranFile -- an independent implementation of the
gfortran record framing the reference's `form="unformatted", access="sequential"` files have.  CPU only."""
import os
import struct

import numpy as np
import pytest
from scipy.io import FortranFile

import mglc_b200 as mg
from mglc_b200 import formats as F
from oracle import formats as OF


def fields(shape, seed, n):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.standard_normal(shape)) for _ in range(n)]


def test_oracle_framing_matches_scipy_fortranfile(tmp_path):
    """the restated framing is what an independent gfortran-format writer produces, and reads back through it"""
    u, v, rho = fields((7, 5, 3), 1, 3)
    p = tmp_path / "scipy.bin"
    with FortranFile(str(p), "w") as ff:
        for a in (u, v, rho):
            ff.write_record(a.T)               # scipy writes C order: the transpose gives column-major element order
    assert p.read_bytes() == OF.output_binary_lid(u, v, rho)
    q = tmp_path / "oracle.bin"
    q.write_bytes(OF.output_binary_lid(u, v, rho))
    with FortranFile(str(q), "r") as ff:
        for a in (u, v, rho):
            assert np.array_equal(ff.read_reals(np.float64).reshape(a.shape, order="F"), a)


def test_output_binary_lid_bytes(tmp_path):
    """L3/output.f90:350-367: three records u, v, rho -- w is not written"""
    u, v, rho = fields((9, 6, 4), 2, 3)
    p = tmp_path / F.output_filename(F.FILE_LID_BIN, 2000)
    assert p.name == "MRTcavity-2000.bin"
    F.output_binary_lid(p, u, v, rho)
    assert p.read_bytes() == OF.output_binary_lid(u, v, rho)
    with FortranFile(str(p), "r") as ff:
        assert np.array_equal(ff.read_reals(np.float64).reshape(u.shape, order="F"), u)


def test_output_binary_thermal_bytes(tmp_path):
    u, v, w, T = fields((5, 6, 7), 3, 4)
    p = tmp_path / F.output_filename(F.FILE_THERMAL_BIN, 31)
    assert p.name == "buoyancyCavity-31.bin"
    F.output_binary_thermal(p, u, v, w, T)
    assert p.read_bytes() == OF.output_binary_thermal(u, v, w, T)


def test_backup_round_trip_and_bytes(tmp_path):
    """backupData() then initial(loadInitField = 1), B3/seq/bouyancy3d.F90:1011-1029, 367-378"""
    n = (6, 5, 4)
    u, v, w, T = fields(n, 4, 4)
    f, g = fields((19,) + n, 5, 1)[0], fields((7,) + n, 6, 1)[0]
    p = tmp_path / F.output_filename(F.FILE_BACKUP, 1000)
    assert p.name == "backupFile-1000.bin"
    F.backup_write(p, u, v, w, T, f, g)
    assert p.read_bytes() == OF.backup_data(u, v, w, T, f, g)
    b = F.backup_read(p, n)
    for k, a in zip(("u", "v", "w", "T", "f", "g"), (u, v, w, T, f, g)):
        assert np.array_equal(b[k], a), k
    with FortranFile(str(p), "r") as ff:                    # an independent reader sees the same six records
        for a in (u, v, w, T, f, g):
            assert np.array_equal(ff.read_reals(np.float64).reshape(a.shape, order="F"), a)
    with pytest.raises(mg.MglcError):                       # wrong grid: record lengths do not match
        F.backup_read(p, (6, 5, 5))


def test_subrecords_of_long_records(tmp_path):
    """records above the subrecord limit (2147483639 bytes in gfortran; 100 here) are split, markers signed"""
    a = np.arange(40, dtype=np.float64)                      # 320 bytes -> 100 + 100 + 100 + 20
    p = tmp_path / "sub.bin"
    F.unformatted_write(p, [a, a[:3]], max_subrecord=100)
    raw = p.read_bytes()
    assert raw == OF.fortran_record(a.tobytes(), 100) + OF.fortran_record(a[:3].tobytes(), 100)
    heads, tails, pos = [], [], 0
    for _ in range(4):
        h = struct.unpack_from("<i", raw, pos)[0]
        t = struct.unpack_from("<i", raw, pos + 4 + abs(h))[0]
        heads.append(h); tails.append(t); pos += 8 + abs(h)
    assert heads == [-100, -100, -100, 20] and tails == [100, -100, -100, -20]
    back = F.unformatted_read(p, [(40,), (3,)])
    assert np.array_equal(back[0], a) and np.array_equal(back[1], a[:3])
    (tmp_path / "cut.bin").write_bytes(raw[:150])
    with pytest.raises(mg.MglcError):
        F.unformatted_read(tmp_path / "cut.bin", [(40,)])


@pytest.mark.parametrize("kind", ["lid", "thermal"])
def test_tecplot_bytes(tmp_path, kind):
    """output_Tecplot(): L3/output.f90:175-313 (Pressure = rho/3) and B3:1623-1773 (T)"""
    n = (6, 4, 5)
    u, v, w, s = fields(n, 7, 4)
    if kind == "lid":
        p = tmp_path / F.output_filename(F.FILE_LID_PLT, 4000)
        assert p.name == "MRTcavity-000004000.plt"
        F.output_tecplot_lid(p, u, v, w, s)
        ref = OF.output_tecplot(u, v, w, s, "Pressure", True)
    else:
        p = tmp_path / F.output_filename(F.FILE_THERMAL_PLT, 12)
        assert p.name == "buoyancyCavity-12.plt"
        F.output_tecplot_thermal(p, u, v, w, s)
        ref = OF.output_tecplot(u, v, w, s, "T", False)
    raw = p.read_bytes()
    assert raw == ref
    assert raw[:8] == b"#!TDV101" and struct.unpack_from("<i", raw, 8)[0] == 1
    data = np.frombuffer(raw[-4 * 7 * np.prod(n):], dtype="<f4").reshape(n[2], n[1], n[0], 7)
    assert data[2, 1, 3, 0] == np.float32(3.5) and data[2, 1, 3, 1] == np.float32(1.5) and data[2, 1, 3, 2] == np.float32(2.5)
    assert data[2, 1, 3, 3] == np.float32(u[3, 1, 2])


def test_get_velocity_and_grid():
    n = (9, 7, 8)
    u, w = fields(n, 8, 2)
    got, ref = F.get_velocity(u, w, 0.1), OF.get_velocity(u, w, 0.1)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert np.array_equal(F.grid_coords(65), OF.grid_coords(65))
    assert F.grid_coords(5).tolist() == [0.0, 0.5, 1.5, 2.5, 3.5, 4.5, 5.0]


def test_format_errors_are_reported_not_thrown(tmp_path):
    u, v, rho = fields((3, 3, 3), 9, 3)
    with pytest.raises(mg.MglcError) as e:
        F.output_binary_lid(tmp_path / "no_such_dir" / "x.bin", u, v, rho)
    assert "cannot open" in str(e.value)
    with pytest.raises(mg.MglcError):
        F.output_filename(99, 1)


def test_thermal2d_files(tmp_path):
    """the 2-D thermal driver's output_binary() (u, v, T) and backupData() (f, g, u, v, T), in the MPI program's layout
    f(0:8,nx,ny) and the OpenACC program's f(nx,ny,0:8) -- mpi_blocked/output.F90:192-217, 381-401; seq/bouyancy2d_acc.F90:1158-1189"""
    n = (7, 5)
    u, v, T = fields(n, 11, 3)
    p = tmp_path / "buoyancyCavity-2000.bin"
    F.output_binary_thermal2d(p, u, v, T)
    assert p.read_bytes() == OF.output_binary_thermal2d(u, v, T)
    with FortranFile(str(p), "r") as ff:
        for a in (u, v, T):
            assert np.array_equal(ff.read_reals(np.float64).reshape(n, order="F"), a)
    for last in (False, True):
        f = fields(n + (9,) if last else (9,) + n, 12, 1)[0]
        g = fields(n + (5,) if last else (5,) + n, 13, 1)[0]
        q = tmp_path / f"backupFile-{int(last)}.bin"
        F.backup_write_2d(q, f, g, u, v, T)
        assert q.read_bytes() == OF.backup_data_2d(f, g, u, v, T)
        b = F.backup_read_2d(q, n, population_last=last)
        for k, a in zip(("f", "g", "u", "v", "T"), (f, g, u, v, T)):
            assert np.array_equal(b[k], a), (last, k)
        with FortranFile(str(q), "r") as ff:
            for a in (f, g, u, v, T):
                assert np.array_equal(ff.read_reals(np.float64).reshape(a.shape, order="F"), a)
    with pytest.raises(mg.MglcError):
        F.backup_read_2d(q, (7, 6))
