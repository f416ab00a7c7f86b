#!/bin/bash
# The GPU calls of round 2, one function per call (so that every file under profiles/r2* can be traced to its command):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh s1'
set -u
S=${1:?session name}
O=gpurun_out/r2$S
mkdir -p $O
NCU="ncu --clock-control none"
clk() { nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader | head -8; }

s1() {   # 1 GPU: host topology, the whole GPU suite, ncu captures of the kernels as they are now
    python tools/probe_host.py --gpus 1 --gib 4 > $O/probe_host.txt 2>&1; tail -4 $O/probe_host.txt
    (timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log); tail -4 $O/pytest_gpu.log
    # dram bytes + duration of the fused lid kernel at the headline size (single pass, no replay)
    timeout 600 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_fused -s 3 -c 2 --csv \
        --log-file $O/ncu_dram_k_fused_768.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/b1.log 2>&1; echo "ncu dram lid rc=$?"
    timeout 600 $NCU --set full --import-source on -k regex:k_fused -s 3 -c 1 -o $O/ncu_full_k_fused_512 \
        python bench.py --size 512 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/b2.log 2>&1; echo "ncu full lid rc=$?"
    timeout 600 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_th_fused -s 3 -c 2 --csv \
        --log-file $O/ncu_dram_k_th_fused_512.csv python bench.py --workload thermal --steps 3 --warmup 3 --no-e2e --no-cpu > $O/b3.log 2>&1; echo "ncu dram thermal rc=$?"
    timeout 600 $NCU --set full --import-source on -k regex:k_th_fused -s 3 -c 1 -o $O/ncu_full_k_th_fused_256 \
        python bench.py --workload thermal --size 256 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/b4.log 2>&1; echo "ncu full thermal rc=$?"
    timeout 600 $NCU --set full --import-source on -k regex:k_jacobi3d -s 5 -c 1 -o $O/ncu_full_k_jacobi3d_512 \
        python bench.py --workload jacobi --steps 10 --warmup 3 --no-e2e --no-cpu > $O/b5.log 2>&1; echo "ncu full jacobi rc=$?"
    timeout 300 python bench.py --workload jacobi --steps 300 --no-cpu > $O/bench_jacobi_512.json 2> $O/b6.err; tail -c 600 $O/bench_jacobi_512.json
    timeout 300 python bench.py --workload thermal --steps 50 --no-cpu > $O/bench_thermal_512.json 2> $O/b7.err; tail -c 600 $O/bench_thermal_512.json
}

"$S"
clk
ls -la $O | tail -30
