# one GPU call: the new 2-D thermal path (tests, smoke, bench, ncu launch list + DRAM traffic), then the whole GPU suite and the headline bench
mkdir -p gpurun_out/s6
O=gpurun_out/s6
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt 2>&1
(timeout 500 python -m pytest tests/test_thermal2d_gpu.py -m gpu -q -x > $O/pytest_t2d.log 2>&1; echo rc=$? >> $O/pytest_t2d.log)
tail -15 $O/pytest_t2d.log
(timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo rc=$? >> $O/pytest_gpu.log)
tail -6 $O/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log); tail -3 $O/smoke.log
(timeout 300 python bench.py --workload thermal2d --steps 50 > $O/bench_thermal2d.json 2> $O/bench_thermal2d.err; echo rc=$?); cat $O/bench_thermal2d.json; tail -3 $O/bench_thermal2d.err
(timeout 200 python bench.py --workload thermal2d --steps 50 --size 16384 --no-cpu > $O/bench_thermal2d_16k.json 2> $O/bench_thermal2d_16k.err; echo rc=$?); cat $O/bench_thermal2d_16k.json
(timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/launches_thermal2d.csv python bench.py --workload thermal2d --steps 6 --warmup 3 --no-cpu > $O/ncu_thermal2d.log 2>&1; echo rc=$?)
(timeout 400 python bench.py --steps 30 > $O/bench_lid.json 2> $O/bench_lid.err; echo rc=$?); cat $O/bench_lid.json
