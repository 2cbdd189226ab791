set -u
mkdir -p gpurun_out
T="timeout -k 5"
$T 200 python scripts/profile_batch.py --batch 4 > gpurun_out/u5_batch_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/u5_batch_plain.log
$T 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"err_kernel|quant_form_y_bf16|scale_den" -s 3 -c 6 -o gpurun_out/u5_elem python scripts/profile_batch.py --batch 4 > gpurun_out/u5_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/u5_ncu.log
