# First GPU call of the next round (one GPU, ~6 min): everything that was written after this round's GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_session_next.sh'
# 1. the parity files that have only been proven on the CPU (host-compiled kernel source): the incompressible and the SRT 2-D
#    lid variants, the plain-C Laplace driver
# 2. the whole GPU suite once (the refactored lid2d_exact.inl is covered by test_lid2d_gpu.py)
# 3. memcheck of one small run of each new variant
mkdir -p gpurun_out/next
O=gpurun_out/next
(timeout 400 python -m pytest tests/test_zz_lid2d_incompressible_gpu.py tests/test_zz_lid2d_srt_gpu.py tests/test_zz_examples_laplace_gpu.py -q -m gpu > $O/zz.log 2>&1; echo zz rc=$?); tail -5 $O/zz.log
(timeout 900 python -m pytest tests -x -q -m gpu > $O/all.log 2>&1; echo all rc=$?); tail -3 $O/all.log
cat > $O/mc.py <<'PY'
import numpy as np, mglc_b200 as mg
for v in ("i", "s"):
    for strict in (True, False):
        s = mg.LidDrivenCavity2D((67, 45), nprocs=4, variant=v, strict=strict, Re=100.0); s.initial(); s.step(70); print(v, strict, s.check()); s.close()
        s = mg.LidDrivenCavity2D((67, 45), variant=v, strict=strict, Re=100.0); s.initial(); s.step(140); print(v, strict, s.check()); s.close()
PY
(timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python $O/mc.py > $O/memcheck.log 2>&1; echo memcheck rc=$?); tail -4 $O/memcheck.log
# 4. bench lines of the new 2-D lid variants (8192^2, HBM-bound like the others)
for v in i s; do (MGLC_BENCH_L2D_VARIANT=$v timeout 200 python bench.py --workload lid2d --steps 50 --warmup 5 --no-cpu > $O/bench_lid2d_8192_variant_$v.json 2> $O/b_$v.err; echo bench $v rc=$?); tail -1 $O/bench_lid2d_8192_variant_$v.json | cut -c1-300; done
