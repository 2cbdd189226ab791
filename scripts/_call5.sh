set -u
mkdir -p gpurun_out
export CB_ENGINE_MAX_BYTES=140000000000
export CB_SCHEDULER_TIMES=0
for cfg in "3 48" "4 64" "5 80" "6 72" "6 96" "4 92" "5 115"; do set -- $cfg
  timeout -k 5 200 python scripts/probe_model_job.py --slots $1 --streams $2 --passes 5 > gpurun_out/u7_model_$1_$2.log 2>&1
  echo "slots $1 streams $2: $(grep -o 'wall [0-9.]* s' gpurun_out/u7_model_$1_$2.log | tr '\n' ' ')"
done
for cfg in "3 48" "5 80" "6 96"; do set -- $cfg
  timeout -k 5 200 python scripts/probe_model_job.py --slots $1 --streams $2 --passes 4 --lbits 4 > gpurun_out/u7_model4_$1_$2.log 2>&1
  echo "lr4 slots $1 streams $2: $(grep -o 'wall [0-9.]* s' gpurun_out/u7_model4_$1_$2.log | tr '\n' ' ')"
done
