set -u
mkdir -p gpurun_out
T="timeout -k 5"
CB_LIBRARY=$PWD/ee274_convexcaldera_llm_quantization_b200/libcaldera_b200_measure.so $T 200 python scripts/probe_err_pass.py > gpurun_out/u6_err_pass.log 2>&1; echo "probe err rc=$?"
cat gpurun_out/u6_err_pass.log
$T 900 python -m pytest tests -x -q -m gpu > gpurun_out/u6_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/u6_pytest.log
export CB_ENGINE_MAX_BYTES=135000000000
$T 400 python bench.py --slots 6 --batch 23 --no-cpu --no-model --no-ref-cuda --no-parity > gpurun_out/u6_bench_6x23.json 2> gpurun_out/u6_bench_6x23.err; echo "bench rc=$?"
python - gpurun_out/u6_bench_6x23.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), {k: round(v['frac'],3) for k,v in d['roofline_all'].items()}, d['clocks'])
PY
