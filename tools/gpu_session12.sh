# one GPU call: the C example driver, a Jacobi register-blocking sweep, compute-sanitizer memcheck over the new kernels
mkdir -p gpurun_out/s12
O=gpurun_out/s12
(timeout 200 python -m pytest tests/test_examples.py -m gpu -q 2>&1 | tail -3)
for jry in 4 8; do for kch in 8 16 32; do
  (MGLC_JACOBI_JRY=$jry MGLC_JACOBI_KCH=$kch timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_jry${jry}_kch${kch}.json 2>> $O/err.txt)
  python -c "
import json;d=json.loads(open('$O/bench_jacobi_jry${jry}_kch${kch}.json').read().strip().splitlines()[-1]);print('jacobi JRY=$jry KCH=$kch', d['value'], d['unit'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
done; done
(MGLC_JACOBI_JRY=8 MGLC_JACOBI_KCH=16 timeout 200 python -m pytest tests/test_jacobi_gpu.py -m gpu -q 2>&1 | tail -2)
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_aa_gpu.py tests/test_thermal2d_gpu.py tests/test_lid2d_gpu.py -m gpu -q -x -k "(strict_is_bit_exact_for_every_way and mrt and calls8) or (acc_fused_step_strict and 23) or (graph_replayed and mpi) or (graph_replayed and c) or (test_fused_step_strict_is_bit_exact and bcT0 and 23 and 6)" > $O/sanitizer_memcheck.log 2>&1; echo sanitizer rc=$?; tail -6 $O/sanitizer_memcheck.log)
tail -3 $O/err.txt
