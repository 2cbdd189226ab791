#!/bin/bash
# Round-2 profiling pass on the GPU box (under gpurun).  Every program first runs plain (must exit 0), then under ncu.
# Outputs land in gpurun_out/r2f_*; scripts/ncu_summary.py / summarize_launches.py turn them into profiles/*.md here.
set -u
mkdir -p gpurun_out
T="timeout -k 5"
# (1) the two roofline kernels alone, full metric set (23 layers per launch = the bench default)
$T 120 python scripts/ncu_targets.py > gpurun_out/r2f_targets_plain.log 2>&1 || { echo "targets plain failed"; exit 1; }
$T 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel|quant_stream_kernel|quant_fast_kernel' -c 6 \
    -o gpurun_out/r2f_targets python scripts/ncu_targets.py > gpurun_out/r2f_targets_ncu.log 2>&1; echo "targets ncu rc=$?"
# (2) launch list of one 23-layer batch through the batched driver
$T 200 python scripts/profile_batch.py --batch 23 > gpurun_out/r2f_batch_plain.log 2>&1 || { echo "batch plain failed"; exit 1; }
$T 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2f_batch_launches.csv python scripts/profile_batch.py --batch 23 > gpurun_out/r2f_batch_ncu.log 2>&1; echo "batch ncu rc=$?"
tail -1 gpurun_out/r2f_batch_plain.log
# (3) launch list of bench.py itself (reduced configuration so that the serialised capture stays short)
BF="--steps 1 --warmup 3 --slots 2 --batch 8 --no-cpu --no-model --no-ref-cuda --no-parity"
$T 300 python bench.py $BF > gpurun_out/r2f_bench_small.json 2> gpurun_out/r2f_bench_small.err || { echo "bench small failed"; exit 1; }
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_bench_launches.csv \
    python bench.py $BF > gpurun_out/r2f_bench_ncu.log 2>&1; echo "bench ncu rc=$?"
