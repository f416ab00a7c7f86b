# one GPU call: CUDA-graph replay of the small 2-D lattices (tests + the shipped sizes' benches)
mkdir -p gpurun_out/s9
O=gpurun_out/s9
(timeout 400 python -m pytest tests/test_lid2d_gpu.py tests/test_thermal2d_gpu.py -m gpu -q > $O/pytest_2d.log 2>&1; echo rc=$? >> $O/pytest_2d.log)
tail -8 $O/pytest_2d.log
(timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log); tail -2 $O/smoke.log
(timeout 200 python bench.py --workload thermal2d --variant acc > $O/bench_thermal2d_acc_513x257.json 2> $O/a.err; echo rc=$?); cat $O/bench_thermal2d_acc_513x257.json; tail -2 $O/a.err
(timeout 200 python bench.py --workload thermal2d --size 201 --steps 4000 > $O/bench_thermal2d_201.json 2> $O/b.err; echo rc=$?); cat $O/bench_thermal2d_201.json; tail -2 $O/b.err
(timeout 200 python bench.py --workload lid2d --size 201 --steps 4000 > $O/bench_lid2d_201.json 2> $O/c.err; echo rc=$?); cat $O/bench_lid2d_201.json; tail -2 $O/c.err
(timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_thermal2d_acc_513x257.csv python bench.py --workload thermal2d --variant acc --steps 130 --no-cpu > $O/ncu_acc.log 2>&1; echo rc=$?)
