# 8 GPUs: BASELINE config 4 (thermal 512^3 = 2x2x2 blocks of 256^3) and config 2 (Jacobi 512^3 per GPU, weak and strong)
mkdir -p gpurun_out/s13
O=gpurun_out/s13
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
(timeout 150 $TR bench.py --gpus 8 --workload thermal --steps 100 --warmup 10 --no-cpu --no-e2e > $O/bench_thermal_512_8gpu.json 2> $O/t.err; echo rc=$?); tail -1 $O/bench_thermal_512_8gpu.json | cut -c1-400
(timeout 100 $TR bench.py --gpus 8 --workload jacobi --steps 300 --warmup 10 --no-cpu --no-e2e > $O/bench_jacobi_weak_8gpu.json 2> $O/j.err; echo rc=$?); tail -1 $O/bench_jacobi_weak_8gpu.json | cut -c1-400
(timeout 100 $TR bench.py --gpus 8 --workload jacobi --scaling strong --steps 300 --warmup 10 --no-cpu --no-e2e > $O/bench_jacobi_strong_8gpu.json 2> $O/js.err; echo rc=$?); tail -1 $O/bench_jacobi_strong_8gpu.json | cut -c1-400
tail -2 $O/t.err $O/j.err
