set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_scheduler.py tests/test_model_driver.py -x -q -m gpu > gpurun_out/u9_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/u9_pytest.log
for st in 1 1 0 1; do
  CB_ENGINE_STAGING=$st timeout -k 5 400 python bench.py --no-cpu --no-model --no-ref-cuda --no-parity > gpurun_out/u9_bench_st$st.json 2> gpurun_out/u9_bench_st$st.err; echo "bench staging=$st rc=$?"
  python - gpurun_out/u9_bench_st$st.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['submit_side_step_seconds'], d['clocks']['sm_mhz'], d['e2e_errors_last_layer']['LR'][:2])
except Exception as e: print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
