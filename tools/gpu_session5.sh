mkdir -p gpurun_out/s5
O=gpurun_out/s5
(timeout 400 python -m pytest tests/test_lid2d_gpu.py tests/test_diagnostics_gpu.py -m gpu -q -x > $O/pytest_new.log 2>&1; echo rc=$? >> $O/pytest_new.log)
tail -15 $O/pytest_new.log
(timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo rc=$? >> $O/pytest_gpu.log)
tail -5 $O/pytest_gpu.log
(timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log); tail -3 $O/smoke.log
(timeout 200 python bench.py --workload lid2d --steps 50 > $O/bench_lid2d.json 2> $O/bench_lid2d.err; echo rc=$?); cat $O/bench_lid2d.json; tail -3 $O/bench_lid2d.err
(timeout 200 python bench.py --workload lid2d --steps 50 --size 16384 --no-cpu > $O/bench_lid2d_16k.json 2> $O/bench_lid2d_16k.err; echo rc=$?); cat $O/bench_lid2d_16k.json
(timeout 250 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/launches_lid2d.csv python bench.py --workload lid2d --steps 6 --warmup 3 --no-cpu > $O/ncu_lid2d.log 2>&1; echo rc=$?)
(timeout 300 python bench.py --steps 30 > $O/bench_lid.json 2> $O/bench_lid.err; echo rc=$?); cat $O/bench_lid.json
