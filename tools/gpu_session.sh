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

m2() {   # 2 GPUs: multi-process parity, the N=2 bench line, which decomposition axis costs what, NVLink traffic of k_fused<0,1>
    (timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 30 $O/pytest_multigpu.log
    (timeout 600 python -m pytest tests/test_jacobi_gpu.py tests/test_particles_gpu.py -q -m gpu -x > $O/pytest_jacobi_particles.log 2>&1; echo "pytest rc=$?" >> $O/pytest_jacobi_particles.log); tail -n 5 $O/pytest_jacobi_particles.log
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_lid_2gpu.json 2> $O/b1.err; tail -c 2500 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err
    for d in 2,1,1 1,2,1 1,1,2; do timeout 300 python tools/group_bench.py --gpus 2 --dims $d --size 768 --steps 20 2>> $O/err.txt | tee -a $O/group_bench_axes.jsonl; done
    timeout 300 python tools/group_bench.py --gpus 1 --dims 1,1,1 --size 768 --steps 20 2>> $O/err.txt | tee -a $O/group_bench_axes.jsonl
    ncu --query-metrics 2>/dev/null | grep -iE "nvl|fabric|pcie" > $O/ncu_metric_names_nvlink.txt; wc -l $O/ncu_metric_names_nvlink.txt
    timeout 900 $NCU --set full --import-source on -k regex:k_fused -s 4 -c 2 -o $O/ncu_full_k_fused_peer_512 \
        python tools/group_bench.py --gpus 2 --dims 2,1,1 --size 512 --steps 2 > $O/b2.log 2>&1; echo "ncu peer rc=$?"
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 > $O/bench_jacobi_2gpu.json 2> $O/b3.err; tail -c 1200 $O/bench_jacobi_2gpu.json; tail -n 3 $O/b3.err
    MGLC_NO_DIRECT=1 timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-parity --no-e2e > $O/bench_jacobi_2gpu_nccl.json 2> $O/b4.err; tail -c 500 $O/bench_jacobi_2gpu_nccl.json
    timeout 600 $TR bench.py --gpus 2 --workload particles --steps 20 --warmup 3 > $O/bench_particles_2gpu.json 2> $O/b5.err; tail -c 1500 $O/bench_particles_2gpu.json; tail -n 3 $O/b5.err
    for cfg in "reg 1" "tma 1" "tma 2"; do
        set -- $cfg
        MGLC_JACOBI_KERNEL=$1 MGLC_JACOBI_TMA_CTAS=$2 timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_$1_$2.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_$1_$2.json').read().strip().splitlines()[-1]);print('jacobi $cfg', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_1gpu.json 2> $O/b10.err; tail -c 600 $O/bench_particles_8192_1gpu.json
    tail -n 5 $O/err.txt
}

jk() {    # 1 GPU: planes per CTA of the register-blocked Jacobi sweep (r1 only tried 8 / 16 / 32)
    for kch in 4 6 8 12; do
        MGLC_JACOBI_KCH=$kch timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_kch$kch.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_kch$kch.json').read().strip().splitlines()[-1]);print('jacobi kch $kch', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
}
san() {   # 1 GPU: compute-sanitizer racecheck + initcheck + memcheck over the fused, single-lattice, particle, Jacobi (TMA) and
          # direct-halo-store paths (P subdomains in one process); logs are kept under profiles/
    SEL='(test_fused_step_strict_is_bit_exact and total0-1) or (test_direct_halo_stores_equal_packed_exchange and 2-None) or (test_halo_push_after_the_update and 2-None) or (test_reinitialising and False) or (test_3d_sweep_kernels_bit_exact and total0 and tma-0-0-0) or (test_3d_sweep_kernels_bit_exact and total0 and reg-0) or (test_fused_steps_with_direct_halo_stores and total2 and direct) or (test_decomposed_run_matches_single_rank_oracle and 2-dims0) or (strict_is_bit_exact_for_every_way and mrt and calls8) or (test_case1_options and opts2 and 4)'
    FILES="tests/test_lid_gpu.py tests/test_jacobi_gpu.py tests/test_particles_gpu.py tests/test_aa_gpu.py tests/test_thermal_gpu.py"
    jk
    for tool in memcheck racecheck initcheck; do
        (timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $FILES -m gpu -q -x -k "$SEL" > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a $O/sanitizer_$tool.log)
        grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -n 4
    done
}

s5() {   # 1 GPU: TMA Jacobi shapes after the row-mapping fix, particle sums/refill list, whole suite
    (timeout 900 python -m pytest tests/test_jacobi_gpu.py tests/test_particles_gpu.py -q -m gpu -x > $O/pytest_jacobi_particles.log 2>&1; echo "pytest rc=$?" >> $O/pytest_jacobi_particles.log); tail -n 6 $O/pytest_jacobi_particles.log
    for cfg in "reg 0" "tma 0" "tma 1" "tma 2"; do
        set -- $cfg
        MGLC_JACOBI_KERNEL=$1 MGLC_JACOBI_TMA_SHAPE=$2 timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_$1_shape$2.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_$1_shape$2.json').read().strip().splitlines()[-1]);print('jacobi $cfg', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    for sh in 0 1; do
    MGLC_JACOBI_TMA_SHAPE=$sh timeout 600 $NCU --set full --import-source on -k regex:k_jacobi3d_tma -s 5 -c 1 -o $O/ncu_full_k_jacobi3d_tma_512_shape$sh \
        python bench.py --workload jacobi --steps 10 --warmup 3 --no-e2e --no-cpu > $O/b5.log 2>&1; echo "ncu full jacobi tma shape $sh rc=$?"
    done
    timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu > $O/bench_particles_8192_1gpu.json 2> $O/b10.err; tail -c 900 $O/bench_particles_8192_1gpu.json; tail -n 3 $O/b10.err
    timeout 300 python bench.py --workload particles --steps 200 --no-cpu > $O/bench_particles_shipped.json 2> $O/b11.err; tail -c 500 $O/bench_particles_shipped.json
    timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 40 -c 24 --csv --log-file $O/launches_particles_8192.csv \
        python bench.py --workload particles --size 8192 --steps 4 --warmup 3 --no-e2e --no-cpu > $O/b12.log 2>&1; echo "ncu particles rc=$?"
    (timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log); tail -n 4 $O/pytest_gpu.log
    tail -n 5 $O/err.txt
}

m3() {   # 2 GPUs: after the fixes -- bench lines (lid with strong leg, jacobi direct vs NCCL, particles), axis costs again
    (timeout 900 python -m pytest tests/test_jacobi_gpu.py tests/test_particles_gpu.py tests/test_lid_gpu.py tests/test_thermal_gpu.py -q -m gpu -x > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log); tail -n 5 $O/pytest_sel.log
    for cfg in "0 8" "1 8" "1 16" "1 32" "0 16"; do
        set -- $cfg
        MGLC_JACOBI_PF=$1 MGLC_JACOBI_KCH=$2 timeout 120 python bench.py --workload jacobi --steps 300 --no-cpu --no-e2e > $O/bench_jacobi_pf$1_kch$2.json 2>> $O/err.txt
        python -c "
import json;d=json.loads(open('$O/bench_jacobi_pf$1_kch$2.json').read().strip().splitlines()[-1]);print('jacobi pf kch $cfg', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 3 $O/pytest_multigpu.log
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_lid_2gpu.json 2> $O/b1.err; tail -c 3000 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err
    for d in 2,1,1 1,1,2; do timeout 300 python tools/group_bench.py --gpus 2 --dims $d --size 768 --steps 20 2>> $O/err.txt | tee -a $O/group_bench_axes.jsonl; done
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 > $O/bench_jacobi_2gpu.json 2> $O/b3.err; tail -c 1200 $O/bench_jacobi_2gpu.json; tail -n 3 $O/b3.err
    MGLC_NO_DIRECT=1 timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-parity --no-e2e > $O/bench_jacobi_2gpu_nccl.json 2> $O/b4.err; tail -c 500 $O/bench_jacobi_2gpu_nccl.json
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --scaling strong --steps 1000 --no-parity --no-e2e > $O/bench_jacobi_2gpu_strong.json 2> $O/b6.err; tail -c 500 $O/bench_jacobi_2gpu_strong.json
    timeout 600 $TR bench.py --gpus 2 --workload particles --steps 20 --warmup 3 > $O/bench_particles_2gpu.json 2> $O/b5.err; tail -c 1500 $O/bench_particles_2gpu.json; tail -n 3 $O/b5.err
    timeout 600 $TR bench.py --gpus 2 --workload thermal --steps 20 --warmup 5 > $O/bench_thermal_2gpu.json 2> $O/b7.err; tail -c 2500 $O/bench_thermal_2gpu.json; tail -n 3 $O/b7.err
    timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_1gpu.json 2> $O/b10.err; tail -c 600 $O/bench_particles_8192_1gpu.json
    tail -n 5 $O/err.txt
}

s6() {   # 1 GPU: where the particle step at scale spends its time; Jacobi default again; the whole suite
    timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 40 -c 24 --csv --log-file $O/launches_particles_8192.csv \
        python bench.py --workload particles --size 8192 --steps 4 --warmup 3 --no-e2e --no-cpu > $O/b12.log 2>&1; echo "ncu particles rc=$?"
    timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_1gpu.json 2> $O/b10.err; tail -c 600 $O/bench_particles_8192_1gpu.json
    timeout 300 python bench.py --workload jacobi --steps 300 --no-cpu > $O/bench_jacobi_512.json 2>> $O/err.txt; tail -c 700 $O/bench_jacobi_512.json
    (timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log); tail -n 4 $O/pytest_gpu.log
    tail -n 5 $O/err.txt
}

m4() {   # 2 GPUs: particle node kernel A/B, lid weak after the barrier-word move, Jacobi direct vs NCCL after the face flag
    (timeout 900 python -m pytest tests/test_particles_gpu.py tests/test_jacobi_gpu.py -q -m gpu -x > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log); tail -n 5 $O/pytest_sel.log
    for v in 1 0; do
        MGLC_P2D_NODE=$v timeout 600 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_1gpu_node$v.json 2> $O/b10.err
        python -c "
import json;d=json.loads(open('$O/bench_particles_8192_1gpu_node$v.json').read().strip().splitlines()[-1]);print('particles node=$v', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 40 -c 24 --csv --log-file $O/launches_particles_8192.csv \
        python bench.py --workload particles --size 8192 --steps 4 --warmup 3 --no-e2e --no-cpu > $O/b12.log 2>&1; echo "ncu particles rc=$?"
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 3 $O/pytest_multigpu.log
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_lid_2gpu.json 2> $O/b1.err; tail -c 1800 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err
    for d in 2,1,1 1,1,2; do timeout 300 python tools/group_bench.py --gpus 2 --dims $d --size 768 --steps 20 2>> $O/err.txt | tee -a $O/group_bench_axes.jsonl; done
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-e2e > $O/bench_jacobi_2gpu.json 2> $O/b3.err; tail -c 700 $O/bench_jacobi_2gpu.json; tail -n 3 $O/b3.err
    MGLC_NO_DIRECT=1 timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-parity --no-e2e > $O/bench_jacobi_2gpu_nccl.json 2> $O/b4.err; tail -c 400 $O/bench_jacobi_2gpu_nccl.json
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --scaling strong --steps 1000 --no-parity --no-e2e > $O/bench_jacobi_2gpu_strong.json 2> $O/b6.err; tail -c 400 $O/bench_jacobi_2gpu_strong.json
    timeout 600 $TR bench.py --gpus 2 --workload particles --steps 20 --warmup 3 --no-e2e > $O/bench_particles_2gpu.json 2> $O/b5.err; tail -c 900 $O/bench_particles_2gpu.json; tail -n 3 $O/b5.err
    timeout 600 $NCU --set full --import-source on -k regex:k_fused -s 4 -c 1 -o $O/ncu_full_k_fused_peer_512 \
        python tools/group_bench.py --gpus 2 --dims 2,1,1 --size 512 --steps 2 > $O/b2.log 2>&1; echo "ncu peer rc=$?"
    tail -n 5 $O/err.txt
}

m5() {   # 2 GPUs: halo push (lid, thermal, Jacobi) against the in-kernel stores and NCCL; case1 particle options; NVLink bytes
    (timeout 900 python -m pytest tests/test_particles_gpu.py tests/test_jacobi_gpu.py tests/test_lid_gpu.py tests/test_thermal_gpu.py -q -m gpu -x > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log); tail -n 5 $O/pytest_sel.log
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 3 $O/pytest_multigpu.log
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_lid_2gpu.json 2> $O/b1.err; tail -c 1500 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-parity --dims 1,1,2 --halo push > $O/bench_lid_2gpu_z_push.json 2> $O/b1z.err; tail -c 400 $O/bench_lid_2gpu_z_push.json
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-parity --dims 1,1,2 --halo direct > $O/bench_lid_2gpu_z_direct.json 2> $O/b1z.err; tail -c 400 $O/bench_lid_2gpu_z_direct.json
    timeout 600 $TR bench.py --gpus 2 --workload thermal --steps 20 --warmup 5 --no-e2e --no-parity > $O/bench_thermal_2gpu.json 2> $O/b7.err; tail -c 900 $O/bench_thermal_2gpu.json
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-e2e > $O/bench_jacobi_2gpu.json 2> $O/b3.err; tail -c 700 $O/bench_jacobi_2gpu.json; tail -n 3 $O/b3.err
    MGLC_NO_DIRECT=1 timeout 600 $TR bench.py --gpus 2 --workload jacobi --steps 300 --no-parity --no-e2e > $O/bench_jacobi_2gpu_nccl.json 2> $O/b4.err; tail -c 400 $O/bench_jacobi_2gpu_nccl.json
    timeout 600 $TR bench.py --gpus 2 --workload jacobi --scaling strong --steps 1000 --no-parity --no-e2e > $O/bench_jacobi_2gpu_strong.json 2> $O/b6.err; tail -c 400 $O/bench_jacobi_2gpu_strong.json
    timeout 600 $NCU --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum -k regex:"k_fused|k_push" -s 6 -c 6 --csv --log-file $O/ncu_nvlink_k_fused_peer_768.csv \
        python tools/group_bench.py --gpus 2 --dims 2,1,1 --size 768 --steps 2 > $O/b2.log 2>&1; echo "ncu nvlink rc=$?"; tail -n 8 $O/ncu_nvlink_k_fused_peer_768.csv | cut -c1-60,200-
    tail -n 5 $O/err.txt 2>/dev/null
}

m6() {   # 2 GPUs: halo push with staged x faces
    (timeout 900 python -m pytest tests/test_lid_gpu.py tests/test_thermal_gpu.py -q -m gpu -x > $O/pytest_sel.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sel.log); tail -n 5 $O/pytest_sel.log
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 3 $O/pytest_multigpu.log
    for d in 2,1,1 1,2,1 1,1,2; do for hm in push direct; do
        timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-parity --dims $d --halo $hm > $O/bench_lid_2gpu_${d//,/}_$hm.json 2> $O/b1z.err
        python -c "
import json;d=json.loads(open('$O/bench_lid_2gpu_${d//,/}_$hm.json').read().strip().splitlines()[-1]);print('lid dims $d $hm', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done; done
    for hm in push direct; do
        timeout 600 $TR bench.py --gpus 2 --workload thermal --steps 20 --warmup 5 --no-e2e --no-parity --dims 2,1,1 --halo $hm > $O/bench_thermal_2gpu_211_$hm.json 2> $O/b7.err
        python -c "
import json;d=json.loads(open('$O/bench_thermal_2gpu_211_$hm.json').read().strip().splitlines()[-1]);print('thermal dims 2,1,1 $hm', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
    timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_lid_2gpu.json 2> $O/b1.err; tail -c 2500 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err
}

m8() {   # 8 GPUs: the driver's own N = 8 line (parity + weak + transports + e2e + strong), then configs 2, 4, 5
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
    python tools/probe_host.py --gpus 0 --gib 1 2>&1 | head -3 > $O/probe_host_8gpu_box.txt
    ( time timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_lid_8gpu.json 2> $O/b1.err ) 2> $O/time_lid.txt; tail -c 3500 $O/bench_lid_8gpu.json; tail -n 3 $O/b1.err; tail -n 3 $O/time_lid.txt
    timeout 300 $TR bench.py --gpus 8 --workload jacobi --steps 300 --no-e2e > $O/bench_jacobi_8gpu_weak.json 2> $O/b3.err; tail -c 600 $O/bench_jacobi_8gpu_weak.json; tail -n 2 $O/b3.err
    timeout 300 $TR bench.py --gpus 8 --workload jacobi --scaling strong --steps 1000 --no-parity --no-e2e > $O/bench_jacobi_8gpu_strong.json 2> $O/b6.err; tail -c 400 $O/bench_jacobi_8gpu_strong.json
    timeout 400 $TR bench.py --gpus 8 --workload thermal --steps 20 --warmup 5 --no-e2e > $O/bench_thermal_8gpu.json 2> $O/b7.err; tail -c 1500 $O/bench_thermal_8gpu.json; tail -n 2 $O/b7.err
    timeout 400 $TR bench.py --gpus 8 --workload particles --steps 10 --warmup 3 --no-e2e > $O/bench_particles_8gpu.json 2> $O/b5.err; tail -c 1000 $O/bench_particles_8gpu.json; tail -n 2 $O/b5.err
}

m9() {   # 2 GPUs: single-lattice (AA) blocks storing into each other; push vs direct at 384^3 blocks
    (timeout 900 python -m pytest tests/test_aa_gpu.py -q -m gpu -x > $O/pytest_aa.log 2>&1; echo "pytest rc=$?" >> $O/pytest_aa.log); tail -n 5 $O/pytest_aa.log
    for d in 2,1,1 1,2,1 1,1,2; do
        timeout 300 python tools/group_bench.py --aa --gpus 2 --dims $d --size 768 --steps 20 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    done
    timeout 300 python tools/group_bench.py --aa --gpus 1 --dims 1,1,1 --size 768 --steps 20 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    timeout 300 python tools/group_bench.py --aa --gpus 2 --dims 1,1,2 --size 896 --steps 10 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    for d in 2,1,1 1,1,2; do for hm in push direct; do
        timeout 300 $TR bench.py --gpus 2 --size 384 --steps 40 --warmup 5 --no-e2e --no-parity --no-cpu --no-extras --dims $d --halo $hm > $O/bench_lid384_2gpu_${d//,/}_$hm.json 2> $O/b1z.err
        python -c "
import json;d=json.loads(open('$O/bench_lid384_2gpu_${d//,/}_$hm.json').read().strip().splitlines()[-1]);print('lid 384 dims $d $hm', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done; done
    tail -n 5 $O/err.txt
}

m10() {   # 2 GPUs: single-lattice blocks, one process per GPU (IPC + flag barrier per launch); NVLink counters of the push
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (MGLC_PARITY_ONLY=lid_aa timeout 300 $TR tests/dist/nccl_worker.py > $O/nccl_worker_lid_aa.log 2>&1; echo "worker rc=$?" >> $O/nccl_worker_lid_aa.log); tail -n 6 $O/nccl_worker_lid_aa.log
    for d in 1,1,2 2,1,1; do
        timeout 300 $TR bench.py --gpus 2 --workload lid_aa --size 768 --steps 20 --warmup 3 --dims $d > $O/bench_lid_aa_2gpu_${d//,/}.json 2> $O/b1.err; tail -c 1800 $O/bench_lid_aa_2gpu_${d//,/}.json; tail -n 3 $O/b1.err
    done
    timeout 300 python tools/group_bench.py --aa --gpus 2 --dims 1,1,2 --size 768 --steps 20 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    timeout 300 python tools/group_bench.py --aa --gpus 2 --dims 2,1,1 --size 768 --steps 20 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    timeout 300 python tools/group_bench.py --aa --gpus 1 --dims 1,1,1 --size 768 --steps 20 >> $O/aa_group_2gpu.jsonl 2>> $O/err.txt; tail -n 1 $O/aa_group_2gpu.jsonl
    ncu --query-metrics 2>/dev/null | grep -i "nvl" | head -40 > $O/ncu_nvlink_metrics.txt; wc -l $O/ncu_nvlink_metrics.txt; head -12 $O/ncu_nvlink_metrics.txt
    MGLC_HALO_MODE=3 timeout 300 $NCU --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:k_push_halos -s 4 -c 4 --csv --log-file $O/ncu_nvlink_push_halos.csv \
        python tools/group_bench.py --gpus 2 --dims 1,1,2 --size 512 --steps 4 > $O/b2.log 2>&1; echo "ncu nvlink rc=$?"; tail -n 6 $O/ncu_nvlink_push_halos.csv
    tail -n 5 $O/err.txt
}

m11() {   # 2 GPUs: single-lattice blocks, one process per GPU, after the block-offset fix
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (MGLC_PARITY_ONLY=lid_aa timeout 300 $TR tests/dist/nccl_worker.py > $O/nccl_worker_lid_aa.log 2>&1; echo "worker rc=$?" >> $O/nccl_worker_lid_aa.log); tail -n 6 $O/nccl_worker_lid_aa.log
    for d in 1,1,2 2,1,1; do
        timeout 300 $TR bench.py --gpus 2 --workload lid_aa --size 768 --steps 20 --warmup 3 --dims $d > $O/bench_lid_aa_2gpu_${d//,/}.json 2> $O/b1.err; tail -c 1800 $O/bench_lid_aa_2gpu_${d//,/}.json; tail -n 3 $O/b1.err
    done
}

s7() {   # 1 GPU: the single-lattice kernels after the per-warp path choice of the odd launch
    (timeout 600 python -m pytest tests/test_aa_gpu.py -q -m gpu -x > $O/pytest_aa.log 2>&1; echo "pytest rc=$?" >> $O/pytest_aa.log); tail -n 3 $O/pytest_aa.log
    timeout 300 python bench.py --workload lid_aa --size 768 --steps 20 --warmup 3 > $O/bench_lid_aa_768.json 2> $O/b1.err; tail -c 900 $O/bench_lid_aa_768.json; tail -n 3 $O/b1.err
    timeout 300 python bench.py --workload lid_aa --steps 20 --warmup 3 > $O/bench_lid_aa_896.json 2> $O/b1.err; tail -c 900 $O/bench_lid_aa_896.json; tail -n 3 $O/b1.err
    timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_aa -c 14 --csv --log-file $O/launches_lid_aa_768.csv \
        python bench.py --workload lid_aa --size 768 --steps 4 --warmup 3 > $O/b2.log 2>&1; echo "ncu launches rc=$?"
    timeout 300 $NCU --set full --import-source on -k regex:k_aa_even -s 1 -c 1 -o $O/ncu_full_k_aa_even_512 python bench.py --workload lid_aa --size 512 --steps 4 --warmup 3 > $O/b3.log 2>&1; echo "ncu even rc=$?"
    timeout 300 $NCU --set full --import-source on -k regex:k_aa_odd -s 1 -c 1 -o $O/ncu_full_k_aa_odd_512 python bench.py --workload lid_aa --size 512 --steps 4 --warmup 3 > $O/b4.log 2>&1; echo "ncu odd rc=$?"
}

s8() {   # 1 GPU: cache hints and occupancy of the single-lattice kernels (MGLC_AA_MEMOP: load = &3, store = >>2&3, 16 / 32 = launch bounds)
    for m in 0 1 2 4 5 8 9 16 32 0; do
        MGLC_AA_MEMOP=$m timeout 200 python bench.py --workload lid_aa --size 640 --steps 100 --warmup 3 > $O/bench_lid_aa_640_memop$m.json 2> $O/b1.err
        python -c "
import json;d=json.loads(open('$O/bench_lid_aa_640_memop$m.json').read().strip().splitlines()[-1]);print('memop $m', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    done
}

m12() {   # 8 GPUs: single-lattice blocks of 960^3 (2x2x2: 1920^3 = 7.08 G cells; the two-lattice path cannot hold a 960^3 block)
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
    ( time timeout 420 $TR bench.py --gpus 8 --workload lid_aa --size 960 --steps 10 --warmup 3 > $O/bench_lid_aa_8gpu_960.json 2> $O/b1.err ) 2> $O/time.txt; tail -c 1900 $O/bench_lid_aa_8gpu_960.json; tail -n 3 $O/b1.err; tail -n 3 $O/time.txt
}
m13() {   # 4 GPUs: parity and the weak-scaling line at N = 4 (1x2x2); e2e / transports / strong left to the round-end run (GPU budget)
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
    ( time timeout 240 $TR bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-extras > $O/bench_lid_4gpu.json 2> $O/b1.err ) 2> $O/time.txt; tail -c 3500 $O/bench_lid_4gpu.json; tail -n 3 $O/b1.err; tail -n 3 $O/time.txt
}

m14() {   # 2 GPUs, round end: the multi-process suite (now with lid_aa) and the default bench line (parity block with the single-lattice verdict)
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
    (timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -x -s > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log); tail -n 4 $O/pytest_multigpu.log
    ( time timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_lid_2gpu.json 2> $O/b1.err ) 2> $O/time.txt; tail -c 2600 $O/bench_lid_2gpu.json; tail -n 3 $O/b1.err; tail -n 3 $O/time.txt
}
s9() {   # 1 GPU, round end: whole GPU suite, smoke, the default bench line
    (timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log); tail -n 4 $O/pytest_gpu.log
    (timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log); tail -n 3 $O/smoke.log
    ( time timeout 600 python bench.py > $O/bench_lid_768_1gpu.json 2> $O/b1.err ) 2> $O/time.txt; tail -c 2000 $O/bench_lid_768_1gpu.json; tail -n 3 $O/b1.err; tail -n 3 $O/time.txt
}

s10() {   # 1 GPU: particle collision kernel with the next row's solid flag loaded ahead (experiment, not adopted: profiles/r2t_*)
    (timeout 400 python -m pytest tests/test_particles_gpu.py -q -m gpu -x > $O/pytest_particles.log 2>&1; echo "pytest rc=$?" >> $O/pytest_particles.log); tail -n 3 $O/pytest_particles.log
    timeout 200 python bench.py --workload particles --size 8192 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_1gpu.json 2> $O/b1.err
    python -c "
import json;d=json.loads(open('$O/bench_particles_8192_1gpu.json').read().strip().splitlines()[-1]);print('particles 8192', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])"
    timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -s 40 -c 24 --csv --log-file $O/launches_particles_8192.csv \
        python bench.py --workload particles --size 8192 --steps 4 --warmup 3 --no-e2e --no-cpu > $O/b12.log 2>&1; echo "ncu particles rc=$?"
}

s11() {   # 1 GPU: A/B of the flag-ahead order in k_p_collision_sum on ONE box (experiment build with the MGLC_P2D_FLAG_AHEAD knob, not adopted)
    for v in 1 0 1 0; do
        MGLC_P2D_FLAG_AHEAD=$v timeout 200 python bench.py --workload particles --size 8192 --steps 30 --warmup 3 --no-cpu --no-e2e > $O/bench_particles_8192_ahead$v.json 2> $O/b1.err
        python -c "
import json;d=json.loads(open('$O/bench_particles_8192_ahead$v.json').read().strip().splitlines()[-1]);print('particles 8192 flag_ahead=$v', d['value'], d['ms_per_step'], 'ms', d['roofline']['frac'])" | tee -a $O/ab.txt
    done
}

san2() {   # 1 GPU: compute-sanitizer memcheck + initcheck + racecheck over what this round added last: single-lattice blocks storing
           # into each other (4 and 12 blocks in one process), the sheared Rayleigh-Benard walls of the 2-D thermal driver
    SEL='(test_group_blocks_strict_bit_exact and mrt and (4-None or 12-dims3)) or test_group_from_initial_and_reinitialised or (test_sheared_walls_fused_step and total1) or (test_sheared_rb_program_as_shipped and 4-dims2)'
    FILES="tests/test_aa_gpu.py tests/test_thermal2d_gpu.py"
    for tool in memcheck initcheck racecheck; do
        (timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $FILES -m gpu -q -x -k "$SEL" > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a $O/sanitizer_$tool.log)
        grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -n 4
    done
}

m15() {   # 8 GPUs: strong scaling of the 768^3 lattice (384^3 blocks) with the shipped default transport (push from 2^25 cells)
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
    timeout 200 $TR bench.py --gpus 8 --scaling strong --size 768 --steps 100 --warmup 10 --no-e2e --no-parity --no-extras --no-cpu > $O/bench_lid_strong_8gpu.json 2> $O/b1.err; tail -c 1500 $O/bench_lid_strong_8gpu.json; tail -n 3 $O/b1.err
}

"$S"
clk
ls -la $O | tail -30
