# one GPU call: 2-D thermal fixes + the single-lattice (AA) path: tests, smoke, benches, ncu launch lists / DRAM traffic / one full capture each
mkdir -p gpurun_out/s7
O=gpurun_out/s7
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > $O/gpu.txt 2>&1
(timeout 300 python -m pytest tests/test_aa_gpu.py tests/test_thermal2d_gpu.py -m gpu -q > $O/pytest_new.log 2>&1; echo rc=$? >> $O/pytest_new.log)
tail -12 $O/pytest_new.log
(timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo rc=$? >> $O/pytest_gpu.log)
tail -6 $O/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log); tail -3 $O/smoke.log
(timeout 300 python bench.py --workload thermal2d --steps 50 > $O/bench_thermal2d.json 2> $O/bench_thermal2d.err; echo rc=$?); cat $O/bench_thermal2d.json; tail -3 $O/bench_thermal2d.err
(timeout 200 python bench.py --workload thermal2d --steps 50 --size 16384 --no-cpu > $O/bench_thermal2d_16k.json 2> $O/bench_thermal2d_16k.err; echo rc=$?); cat $O/bench_thermal2d_16k.json
(timeout 300 python bench.py --workload lid_aa --steps 30 > $O/bench_lid_aa_896.json 2> $O/bench_lid_aa_896.err; echo rc=$?); cat $O/bench_lid_aa_896.json; tail -3 $O/bench_lid_aa_896.err
(timeout 300 python bench.py --workload lid_aa --steps 30 --size 768 > $O/bench_lid_aa_768.json 2> $O/bench_lid_aa_768.err; echo rc=$?); cat $O/bench_lid_aa_768.json
(timeout 300 python bench.py --workload lid_aa --steps 20 --size 960 > $O/bench_lid_aa_960.json 2> $O/bench_lid_aa_960.err; echo rc=$?); cat $O/bench_lid_aa_960.json; tail -2 $O/bench_lid_aa_960.err
(timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O/launches_thermal2d.csv python bench.py --workload thermal2d --steps 6 --warmup 3 --no-cpu > $O/ncu_thermal2d.log 2>&1; echo rc=$?)
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O/launches_lid_aa_768.csv python bench.py --workload lid_aa --size 768 --steps 6 --warmup 3 > $O/ncu_lid_aa.log 2>&1; echo rc=$?)
(timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_aa_odd -c 1 -o $O/full_k_aa_odd_384 python bench.py --workload lid_aa --size 384 --steps 4 --warmup 3 > $O/ncu_full_aa_odd.log 2>&1; echo rc=$?)
(timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_aa_even -c 1 -o $O/full_k_aa_even_384 python bench.py --workload lid_aa --size 384 --steps 4 --warmup 3 > $O/ncu_full_aa_even.log 2>&1; echo rc=$?)
(timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_t2_fused -c 1 -o $O/full_k_t2_fused_4096 python bench.py --workload thermal2d --size 4096 --steps 4 --warmup 3 --no-cpu > $O/ncu_full_t2.log 2>&1; echo rc=$?)
ls -la $O
