"""CPU-side checks of bench.py's contract: the reference arm (`--impl reference`, the CPU restatement timed on the host cores)
prints ONE JSON line with the driver's keys on rank 0 and nothing on the other ranks; the product arm refuses to run without a
CUDA device instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = bench("--impl", "reference", "--size", "32", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "MLUPS" and j["unit"] == "MLUPS" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 1 and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["vs_baseline"] is None and j["value"] > 0 and j["ms_per_step"] > 0
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_every_core_under_a_launcher_and_names_the_product_arms_config():
    """torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arm must still run on all the cores it reports, and
    its `config` must be the N-GPU product arm's (same keys: workload, global_lattice, decomposition, ...)."""
    import bench as B
    r = bench("--impl", "reference", "--gpus", "8", "--size", "32", "--steps", "2", "--warmup", "1",
              env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "8", "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == j["detail"]["omp_threads"]
    assert j["config"]["global_lattice"] == [64, 64, 64] and j["config"]["decomposition"] == "2x2x2" and j["n_gpus"] == 8

    class A:
        size, scaling, dims = 32, "weak", ""
    dims, per_gpu, gn = B.lattice_for(A, 8)
    assert j["config"] == B.lattice_config(A, False, dims, per_gpu, gn)
    assert [B.dims_create(n) for n in (1, 2, 4, 8, 12, 6)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 2), (3, 2, 1)]
    r = bench("--impl", "reference", "--workload", "jacobi", "--gpus", "4", "--size", "24", "--steps", "2", "--warmup", "1", env={"OMP_NUM_THREADS": "1"})
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["impl"] == "reference" and j["metric"] == "Mcells/s" and j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert j["config"]["decomposition"] == "1x2x2" and j["config"]["global_grid"] == [24, 48, 48]


def test_reference_arm_is_silent_on_other_ranks():
    r = bench("--impl", "reference", "--gpus", "2", "--size", "32", "--steps", "2", "--warmup", "1",
              env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not r.stdout.strip()


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = bench("--size", "32", "--steps", "2", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
