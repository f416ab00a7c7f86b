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

s2() {   # 1 GPU: the TMA Jacobi pipeline: parity, tuning sweep, ncu; the reworked bench line at N=1
    (timeout 600 python -m pytest tests/test_jacobi_gpu.py tests/test_lid_gpu.py -q -m gpu -x > $O/pytest_jacobi_lid.log 2>&1; echo "pytest rc=$?" >> $O/pytest_jacobi_lid.log); tail -4 $O/pytest_jacobi_lid.log
    for cfg in "reg 1 64" "tma 1 64" "tma 2 64" "tma 1 32" "tma 1 128" "tma 2 128" "tma 1 512" "tma 2 512" "tma 2 16"; do
        set -- $cfg
        MGLC_JACOBI_KERNEL=$1 MGLC_JACOBI_TMA_CTAS=$2 MGLC_JACOBI_SLAB=$3 timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_$1_$2_$3.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_$1_$2_$3.json').read().strip().splitlines()[-1]);print('jacobi $cfg', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    timeout 600 $NCU --set full --import-source on -k regex:k_jacobi3d_tma -s 5 -c 1 -o $O/ncu_full_k_jacobi3d_tma_512 \
        python bench.py --workload jacobi --steps 10 --warmup 3 --no-e2e --no-cpu > $O/b5.log 2>&1; echo "ncu full jacobi tma rc=$?"
    timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_lid_768_1gpu.json 2> $O/b8.err; tail -c 1500 $O/bench_lid_768_1gpu.json
    timeout 600 python bench.py --workload thermal --steps 20 --warmup 5 > $O/bench_thermal_512_1gpu.json 2> $O/b9.err; tail -c 1200 $O/bench_thermal_512_1gpu.json
    tail -n 5 $O/err.txt $O/b8.err $O/b9.err
}

s3() {   # 1 GPU: triage of the TMA pipeline, then the s2 program again
    timeout 900 python tools/debug_tma.py --sanitize > $O/debug_tma.txt 2>&1; head -c 5000 $O/debug_tma.txt
    if grep -q "{} -> \['bit-exact'\]" $O/debug_tma.txt; then s2; fi
}

s4() {   # 1 GPU: TMA Jacobi L2-promotion sweep, particle bins parity + config 5 at scale on one GPU
    for cfg in "reg 1 0" "tma 1 0" "tma 2 0" "tma 1 128" "tma 1 256" "tma 2 128"; do
        set -- $cfg
        MGLC_JACOBI_KERNEL=$1 MGLC_JACOBI_TMA_CTAS=$2 MGLC_JACOBI_TMA_L2PROMO=$3 timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_$1_$2_promo$3.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_$1_$2_promo$3.json').read().strip().splitlines()[-1]);print('jacobi $cfg', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    timeout 600 $NCU --set full --import-source on -k regex:k_jacobi3d_tma -s 5 -c 1 -o $O/ncu_full_k_jacobi3d_tma_512 \
        python bench.py --workload jacobi --steps 10 --warmup 3 --no-e2e --no-cpu > $O/b5.log 2>&1; echo "ncu full jacobi tma rc=$?"
    (timeout 900 python -m pytest tests/test_particles_gpu.py tests/test_jacobi_gpu.py -q -m gpu -x > $O/pytest_particles.log 2>&1; echo "pytest rc=$?" >> $O/pytest_particles.log); tail -n 6 $O/pytest_particles.log
    timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 > $O/bench_particles_8192_1gpu.json 2> $O/b10.err; tail -c 1500 $O/bench_particles_8192_1gpu.json; tail -n 3 $O/b10.err
    timeout 300 python bench.py --workload particles --steps 200 > $O/bench_particles_shipped.json 2> $O/b11.err; tail -c 700 $O/bench_particles_shipped.json
    timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 60 -c 40 --csv --log-file $O/launches_particles_8192.csv \
        python bench.py --workload particles --size 8192 --steps 5 --warmup 3 --no-e2e --no-cpu > $O/b12.log 2>&1; echo "ncu particles rc=$?"
    (timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log); tail -n 4 $O/pytest_gpu.log
    tail -n 5 $O/err.txt
}

"$S"
clk
ls -la $O | tail -30
