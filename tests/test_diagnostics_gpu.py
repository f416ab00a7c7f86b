"""GPU tests of the in-loop diagnostics and output paths: calNuRe() volume averages reduced on the
device (B3/mpi_blocked/RaNu.F90:13-47), centre-line extraction without a full-field download (getVelocity,
L3/output.f90:318-347), output() files written from device state, and backupData() -> loadInitField restart
(B3/seq/bouyancy3d.F90:1011-1029, 367-378) continuing bit for bit."""
import numpy as np
import pytest

import mglc_b200 as mg
from mglc_b200 import formats as F
from oracle import formats as OF
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nprocs", [1, 2, 8])
def test_calNuRe_matches_oracle(nprocs):
    total = (17, 15, 13)
    wd = orc.ThermalWorld(total, nprocs)
    sim = mg.BuoyancyDrivenCavity(total, nprocs=nprocs, arith="strict")
    wd.initial(); sim.initial()
    nu0, re0 = sim.calNuRe()
    assert nu0 == 1.0 and re0 == 0.0            # fluid at rest: pure conduction
    wd.step(40); sim.step(40)
    (nu, re), (onu, ore) = sim.calNuRe(), wd.calNuRe()
    # the fields are bit-identical (strict); the sums differ only by summation order (tree vs loop order)
    assert np.isclose(nu, onu, rtol=1e-12, atol=0) and np.isclose(re, ore, rtol=1e-12, atol=0), (nu, onu, re, ore)
    # definition check against the gathered fields
    m = sim.gather_macro()
    visc = (sim.tauf - 0.5) / 3.0
    n = float(np.prod(total))
    assert np.isclose(nu, (m["w"] * m["T"]).sum() / n * total[2] / (visc / 0.71) + 1.0, rtol=1e-12)
    assert np.isclose(re, np.sqrt((m["u"] ** 2 + m["v"] ** 2 + m["w"] ** 2).sum() / n) * total[2] / visc, rtol=1e-12)
    sim.step(3)                                  # the diagnostic leaves the fused loop and re-enters it cleanly
    wd.step(3)
    for k in ("rho", "u", "v", "w", "T"):
        assert np.array_equal(sim.gather(k), wd.gather(k)), k
    wd.close(); sim.close()


@pytest.mark.parametrize("nprocs", [1, 8, 12])
def test_getVelocity_lines_without_full_download(nprocs):
    total = (21, 14, 18)
    wd = orc.LidWorld(total, 1)
    sim = mg.LidDrivenCavity(total, nprocs=nprocs, arith="strict")
    wd.initial(); sim.initial(); wd.step(30); sim.step(30)
    u, w = wd.gather("u"), wd.gather("w")
    got, ref = sim.getVelocity(), OF.get_velocity(u, w, 0.1)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert np.array_equal(sim.download_line("rho", 1, 5, 7), wd.gather("rho")[4, :, 6])
    wd.close(); sim.close()


def test_output_files_from_device_state(tmp_path):
    total = (12, 10, 9)
    wd = orc.LidWorld(total, 1)
    sim = mg.LidDrivenCavity(total, nprocs=4, arith="strict")
    wd.initial(); sim.initial(); wd.step(7); sim.step(7)
    plt, binf = sim.output(str(tmp_path), 7)
    assert plt.endswith("MRTcavity-000000007.plt") and binf.endswith("MRTcavity-7.bin")
    g = {k: wd.gather(k) for k in ("rho", "u", "v", "w")}
    assert open(plt, "rb").read() == OF.output_tecplot(g["u"], g["v"], g["w"], g["rho"], "Pressure", True)
    assert open(binf, "rb").read() == OF.output_binary_lid(g["u"], g["v"], g["rho"])
    wd.close(); sim.close()
    wd = orc.ThermalWorld(total, 1)
    sim = mg.BuoyancyDrivenCavity(total, nprocs=2, arith="strict")
    wd.initial(); sim.initial(); wd.step(5); sim.step(5)
    binf, plt = sim.output(str(tmp_path), 5)
    g = {k: wd.gather(k) for k in ("u", "v", "w", "T")}
    assert open(binf, "rb").read() == OF.output_binary_thermal(g["u"], g["v"], g["w"], g["T"])
    assert open(plt, "rb").read() == OF.output_tecplot(g["u"], g["v"], g["w"], g["T"], "T", False)
    wd.close(); sim.close()


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_backup_restart_continues_bit_for_bit(tmp_path, arith):
    """run 12 steps, backupData(), reload into a fresh (differently decomposed) simulation, run 9 more:
    identical to 21 uninterrupted steps"""
    total = (14, 11, 10)
    a = mg.BuoyancyDrivenCavity(total, nprocs=1, arith=arith)
    a.initial(); a.step(12)
    path = a.backupData(str(tmp_path), 12)
    assert path.endswith("backupFile-12.bin")
    wd = orc.ThermalWorld(total, 1)
    wd.initial(); wd.step(12)
    if arith == "strict":
        assert open(path, "rb").read() == OF.backup_data(*[wd.gather(k) for k in ("u", "v", "w", "T", "f", "g")])
    a.step(9)
    b = mg.BuoyancyDrivenCavity(total, nprocs=4, arith=arith)
    b.initial()
    b.loadInitField(path, rho="macro")
    b.step(9)
    for k in ("rho", "u", "v", "w", "T", "f", "g"):
        assert np.array_equal(a.gather(k), b.gather(k)), k
    if arith == "strict":
        # the reference's own restart: rho = 1 until the first macro() (seq:289) -- against the oracle doing the same
        b.loadInitField(path)
        wd.scatter("rho", np.ones(total, order="F"))
        b.step(9); wd.step(9)
        for k in ("rho", "u", "v", "w", "T"):
            assert np.array_equal(b.gather(k), wd.gather(k)), k
    a.close(); b.close(); wd.close()
