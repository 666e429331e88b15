#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU, AFTER the same commands exited 0 without ncu):
#   launch list of the bench command + ncu --set full captures of the dominant kernels, summarised
#   on the GPU box by tools/ncu_summary.py (the .ncu-rep files are too large to bring back whole).
set -u
R=${1:-r02}
mkdir -p /tmp/rep gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/${R}_ncu_bench.log 2>&1
cap() {  # name, kernel regex, count, import-source(on/off), prof_search args...
    local name=$1 rx=$2 cnt=$3 src=$4; shift 4
    local extra=""; [ "$src" = on ] && extra="--import-source on"
    ncu --set full --clock-control none $extra -k "regex:$rx" -c $cnt -o /tmp/rep/${R}_$name \
        python tools/prof_search.py "$@" > gpurun_out/${R}_ncu_$name.log 2>&1
    python tools/ncu_summary.py /tmp/rep/${R}_$name.ncu-rep gpurun_out/${R}_$name
}
cap gemm_b1024 "gemm_topk_kernel|select_warp_kernel" 4 on --rows 1000000 --dim 768 --batch 1024 --iters 1
cap gemm_b1 "gemm_topk_kernel|select_kernel" 4 on --rows 1000000 --dim 768 --batch 1 --iters 1
cap scan_mq "scan_float_mq_kernel" 1 off --rows 1000000 --dim 1536 --batch 8 --metric manhattan --iters 1
cap gemm_cfg3 "gemm_topk_kernel" 2 on --rows 10000000 --dim 128 --batch 4096 --k 100 --metric euclidean --iters 1
cap scan_u8 "scan_quant" 1 off --rows 12500000 --dim 96 --dtype u8 --batch 1 --iters 1
ls -la gpurun_out | tail -30
