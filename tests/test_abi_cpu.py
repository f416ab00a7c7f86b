"""CPU-only checks of the drop-in boundary: libmglc.so loads, exports every symbol include/mglc.h
declares, its host-side decomposition / halo plan agree with the oracle's restatement of
L3/main.f90:24-72,144-212, and the device entry points fail loudly (no CPU fallback) without a GPU.
Also checks the product's per-cell arithmetic source (compiled for the host by tests/host_shim)
against the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import mglc_b200 as mg
from mglc_b200 import _lib as L
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mglc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mglc_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 45
    lib = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mglc.h but not exported by libmglc.so"
    assert set(names) == set(L.SIGNATURES), set(names) ^ set(L.SIGNATURES)
    assert mg.lib().mglc_version() == 105


def test_no_cpu_fallback_without_gpu():
    n = C.c_int()
    mg.lib().mglc_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mg.MglcError) as e:
        mg.LidDrivenCavity((8, 8, 8))
    assert e.value.code == L.E_NOGPU
    with pytest.raises(mg.MglcError) as e:
        mg.LidDrivenCavity((8, 8, 8), nprocs=2)
    assert e.value.code == L.E_NOGPU


def test_invalid_descriptors_are_rejected():
    lib = mg.lib()
    d = mg.make_desc((8, 8, 8))
    h = C.c_void_p()
    d.tau = 0.5
    assert lib.mglc_lbm_create(C.byref(h), C.byref(d), None) == L.E_INVALID
    d = mg.make_desc((8, 8, 8))
    d.ln[0] = 9
    assert lib.mglc_lbm_create(C.byref(h), C.byref(d), None) == L.E_INVALID
    d = mg.make_desc((8, 8, 8), 2, 0)
    assert lib.mglc_lbm_create(C.byref(h), C.byref(d), None) == L.E_INVALID      # needs a communicator
    assert b"communicator" in lib.mglc_last_error()
    assert lib.mglc_decompose_1d(10, 3, 2, C.byref(C.c_int()), None) == L.E_INVALID
    with pytest.raises(mg.MglcError):
        mg.make_desc((2, 8, 8), 4, 0, dims=(4, 1, 1))       # fewer cells than ranks


@pytest.mark.parametrize("nprocs", [1, 2, 3, 4, 6, 8, 12, 16, 27])
def test_decomposition_matches_oracle(nprocs):
    total = (65, 33, 29)
    wd = orc.LidWorld(total, nprocs)
    assert mg.dims_create(nprocs) == wd.dims
    for r, R in enumerate(wd.ranks):
        d = mg.make_desc(total, nprocs, r)
        assert tuple(d.coords) == R.coords and tuple(d.ln) == R.n and tuple(d.start) == R.start
        ns, nl = mg.cart_neighbors(wd.dims, R.coords)
        assert ns == R.nbr_surface and nl == R.nbr_line
        assert d.tau == wd.tauf
    wd.close()


@pytest.mark.parametrize("nprocs,dims", [(2, None), (8, None), (12, None), (4, (1, 4, 1)), (18, (3, 3, 2))])
def test_halo_plan_is_pairwise_consistent(nprocs, dims):
    total = (21, 17, 19)
    plans = [mg.halo_plan(mg.make_desc(total, nprocs, r, dims)) for r in range(nprocs)]
    order = [0, 1, 2, 3, 4, 5, 7, 10, 9, 8, 11, 14, 13, 12, 15, 18, 17, 16]     # ex_sendrecv.f90 call order
    for r, plan in enumerate(plans):
        assert [m["dir"] for m in plan] == order
        pairs = set()
        for q, m in enumerate(plan):
            if m["send_to"] >= 0:
                peer = plans[m["send_to"]][q]
                assert peer["recv_from"] == r and peer["recv_count"] == m["send_count"] > 0
                assert m["send_to"] not in pairs            # at most one message per ordered pair
                pairs.add(m["send_to"])
            else:
                assert m["send_count"] == 0
            if m["recv_from"] >= 0:
                peer = plans[m["recv_from"]][q]
                assert peer["send_to"] == r and peer["send_count"] == m["recv_count"]
    # volume check against the survey's formula: faces carry 5 populations, edges 1
    d = mg.make_desc((768, 768, 768), 8, 0)
    p = mg.halo_plan(d)
    assert sum(m["send_count"] for m in p) == 3 * 5 * 384 * 384 + 3 * 384


def test_relaxation_rates():
    a, b = C.c_double(), C.c_double()
    tau = 0.1 * 65.0 / 1000.0 * 3.0 + 0.5
    assert mg.lib().mglc_relaxation_rates(tau, C.byref(a), C.byref(b)) == 0
    assert a.value == 1.0 / tau and b.value == 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "mrt_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out,
                           os.path.join(ROOT, "tests", "host_shim", "mrt_host.cpp")])
    S = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    S.shim_collide.argtypes = [C.c_int, dp] + [C.c_double] * 6 + [dp]
    S.shim_macro.argtypes = [dp, dp]
    S.shim_collide_bgk.argtypes = [C.c_int, dp] + [C.c_double] * 5 + [dp]
    return S


def test_kernel_arithmetic_source_matches_oracle(shim):
    """mglc_b200/csrc/d3q19_mrt.inl compiled for the host: strict == oracle bitwise, fast to rounding."""
    dp = C.POINTER(C.c_double)
    O = orc.lib()
    rng = np.random.default_rng(42)
    worst = 0.0
    for _ in range(3000):
        rho = 1 + 0.05 * rng.uniform(-1, 1)
        u, v, w = 0.1 * rng.uniform(-1, 1, 3)
        f = np.ascontiguousarray(orc.feq(rho, u, v, w) * (1 + 0.05 * rng.uniform(-1, 1, 19)))
        a, b, c, m = np.zeros(19), np.zeros(19), np.zeros(19), np.zeros(4)
        O.orc_collide_cell(f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, 1.3, a.ctypes.data_as(dp))
        shim.shim_collide(1, f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, 1.3, b.ctypes.data_as(dp))
        shim.shim_collide(0, f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, 1.3, c.ctypes.data_as(dp))
        assert np.array_equal(a, b)
        worst = max(worst, np.abs(a - c).max())
        shim.shim_macro(f.ctypes.data_as(dp), m.ctypes.data_as(dp))
        r = 0.0
        for q in range(19):
            r = r + f[q]
        su = sum_in_order(f, orc.EX)
        assert m[0] == r and m[1] == su / r
    assert worst < 1e-15


def test_bgk_arithmetic_source_matches_oracle(shim):
    """d3q19_collide_bgk (L3/collision.f90:191-198) compiled for the host: strict == oracle bitwise, fast to rounding."""
    dp = C.POINTER(C.c_double)
    O = orc.lib()
    rng = np.random.default_rng(43)
    worst = 0.0
    for _ in range(3000):
        rho = 1 + 0.05 * rng.uniform(-1, 1)
        u, v, w = 0.1 * rng.uniform(-1, 1, 3)
        f = np.ascontiguousarray(orc.feq(rho, u, v, w) * (1 + 0.05 * rng.uniform(-1, 1, 19)))
        a, b, c = np.zeros(19), np.zeros(19), np.zeros(19)
        O.orc_collide_cell_bgk(f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, a.ctypes.data_as(dp))
        shim.shim_collide_bgk(1, f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, b.ctypes.data_as(dp))
        shim.shim_collide_bgk(0, f.ctypes.data_as(dp), rho, u, v, w, 1 / 0.5195, c.ctypes.data_as(dp))
        assert np.array_equal(a, b)
        worst = max(worst, np.abs(a - c).max())
    assert worst < 1e-15


def sum_in_order(f, e):
    s = 0.0
    for q in range(19):
        s = s + f[q] * float(e[q])
    return s
