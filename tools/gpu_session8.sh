# one GPU call: the OpenACC-program variant of the 2-D thermal path, the reworked single-lattice kernels, the whole suite, benches
mkdir -p gpurun_out/s8
O=gpurun_out/s8
(timeout 400 python -m pytest tests/test_aa_gpu.py tests/test_thermal2d_gpu.py -m gpu -q > $O/pytest_new.log 2>&1; echo rc=$? >> $O/pytest_new.log)
tail -12 $O/pytest_new.log
(timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo rc=$? >> $O/pytest_gpu.log)
tail -6 $O/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log); tail -3 $O/smoke.log
(timeout 300 python bench.py --workload lid_aa --steps 30 > $O/bench_lid_aa_896.json 2> $O/bench_lid_aa_896.err; echo rc=$?); cat $O/bench_lid_aa_896.json; tail -3 $O/bench_lid_aa_896.err
(timeout 300 python bench.py --workload lid_aa --steps 30 --size 768 > $O/bench_lid_aa_768.json 2> $O/bench_lid_aa_768.err; echo rc=$?); cat $O/bench_lid_aa_768.json
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O/launches_lid_aa_768.csv python bench.py --workload lid_aa --size 768 --steps 6 --warmup 3 > $O/ncu_lid_aa.log 2>&1; echo rc=$?)
(timeout 300 python bench.py --workload thermal2d --variant acc > $O/bench_thermal2d_acc_513x257.json 2> $O/bench_thermal2d_acc.err; echo rc=$?); cat $O/bench_thermal2d_acc_513x257.json; tail -3 $O/bench_thermal2d_acc.err
(timeout 300 python bench.py --workload thermal2d --variant acc --size 8192 --steps 50 --no-cpu > $O/bench_thermal2d_acc_16385x8193.json 2> $O/bench_thermal2d_acc_big.err; echo rc=$?); cat $O/bench_thermal2d_acc_16385x8193.json; tail -3 $O/bench_thermal2d_acc_big.err
(timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_aa_odd -c 1 -o $O/full_k_aa_odd_384_v2 python bench.py --workload lid_aa --size 384 --steps 4 --warmup 3 > $O/ncu_full_aa_odd.log 2>&1; echo rc=$?)
(timeout 400 python bench.py --steps 30 > $O/bench_lid.json 2> $O/bench_lid.err; echo rc=$?); cat $O/bench_lid.json
