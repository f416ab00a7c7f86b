# A/B: CUDA-graph replay on / off for the shipped small 2-D lattices
mkdir -p gpurun_out/s10
O=gpurun_out/s10
for w in lid2d thermal2d; do for g in 0 65536; do
  (MGLC_2D_GRAPH_CELLS=$g timeout 200 python bench.py --workload $w --size 201 --steps 8000 --no-cpu > $O/bench_${w}_201_graph$g.json 2> $O/err.txt; echo rc=$?); python -c "
import json;d=json.loads(open('$O/bench_${w}_201_graph$g.json').read().strip().splitlines()[-1]);print('$w graph_cells=$g', d['value'], 'MLUPS', d['ms_per_step']*1e3, 'us/step')"
done; done
(MGLC_2D_GRAPH_CELLS=0 timeout 200 python bench.py --workload thermal2d --variant acc --steps 8000 --no-cpu > $O/bench_acc_graph0.json 2>> $O/err.txt); (MGLC_2D_GRAPH_CELLS=1000000 timeout 200 python bench.py --workload thermal2d --variant acc --steps 8000 --no-cpu > $O/bench_acc_graph1.json 2>> $O/err.txt)
python -c "
import json
for g in (0,1):
    d=json.loads(open('$O/bench_acc_graph%d.json'%g).read().strip().splitlines()[-1]);print('acc 513x257 graph=%d'%g, d['value'], 'MLUPS', d['ms_per_step']*1e3, 'us/step')"
(timeout 100 python -m pytest tests/test_lid2d_gpu.py tests/test_thermal2d_gpu.py -m gpu -q -k graph 2>&1 | tail -2)
tail -3 $O/err.txt
