#!/bin/bash
# Runs on the GPU box (under gpurun): plain runs first, then the ncu passes of the profiling recipe.
# Outputs land in gpurun_out/prof_*; scripts/make_profiles.py turns them into profiles/*.md here.
set -u
mkdir -p gpurun_out
T="timeout -k 5"
$T 120 python scripts/profile_layer.py > gpurun_out/prof_layer_plain.log 2>&1 || exit 1
$T 120 python scripts/profile_layer.py --lbits 4 > gpurun_out/prof_layer_lr4_plain.log 2>&1 || exit 1
$T 120 python scripts/profile_layer.py --mode latency > gpurun_out/prof_layer_latency_plain.log 2>&1 || exit 1
$T 60 python scripts/profile_quant.py 2 64 > gpurun_out/prof_quant_plain.log 2>&1 || exit 1
$T 60 python scripts/profile_smalldense.py 224 > gpurun_out/prof_small_plain.log 2>&1 || exit 1
# (1) every launch of one layer with its device time
$T 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/prof_layer_launches.csv python scripts/profile_layer.py > gpurun_out/prof_ncu1.log 2>&1
$T 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/prof_layer_lr4_launches.csv python scripts/profile_layer.py --lbits 4 > gpurun_out/prof_ncu1b.log 2>&1
$T 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/prof_layer_latency_launches.csv python scripts/profile_layer.py --mode latency > gpurun_out/prof_ncu1c.log 2>&1
# (2) full captures of the top kernels
$T 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"gemm_tc_kernel|splitk_reduce|form_y_bf16_kernel|quant_err_kernel|err_kernel|resid_absmax" -c 14 \
    -o gpurun_out/prof_layer python scripts/profile_layer.py > gpurun_out/prof_ncu2.log 2>&1
$T 300 ncu --set full --clock-control none --import-source on -k regex:quant_fast_kernel -s 8 -c 1 \
    -o gpurun_out/prof_quant python scripts/profile_quant.py 2 64 > gpurun_out/prof_ncu3.log 2>&1
$T 300 ncu --set full --clock-control none --import-source on -k regex:"chol_inv_kernel|jacobi_cluster_kernel|jacobi_smem_kernel" -s 4 -c 2 \
    -o gpurun_out/prof_small python scripts/profile_smalldense.py 224 > gpurun_out/prof_ncu4.log 2>&1
cat gpurun_out/prof_layer_plain.log gpurun_out/prof_layer_lr4_plain.log gpurun_out/prof_layer_latency_plain.log gpurun_out/prof_quant_plain.log
tail -2 gpurun_out/prof_small_plain.log
