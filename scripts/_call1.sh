set -u
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python -m pytest tests/test_gpu_quantizer.py -x -q -m gpu > gpurun_out/u1_pytest_quant.log 2>&1; echo "pytest quant rc=$?"
tail -3 gpurun_out/u1_pytest_quant.log
CB_LIBRARY=$PWD/ee274_convexcaldera_llm_quantization_b200/libcaldera_b200_measure.so $T 200 python scripts/probe_quant_stream.py > gpurun_out/u1_quant_stream.log 2>&1; echo "probe quant rc=$?"
grep -E "4096x4096_b2_bs64|8192|mismatch" gpurun_out/u1_quant_stream.log
$T 200 python scripts/probe_gemm2.py 18 23 24 27 32 37 > gpurun_out/u1_gemm2.log 2>&1; echo "gemm2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/u1_gemm2.log'))
for k,v in d.items(): print(k, round(v['tflops'],1), round(v['per_layer_us'],2))
PY
export CB_ENGINE_MAX_BYTES=135000000000
for cfg in "5 23" "6 23" "7 18"; do set -- $cfg
  $T 400 python bench.py --slots $1 --batch $2 --no-cpu --no-model --no-ref-cuda --no-parity > gpurun_out/u1_bench_$1x$2.json 2> gpurun_out/u1_bench_$1x$2.err; echo "bench $1x$2 rc=$?"
  python - gpurun_out/u1_bench_$1x$2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['roofline_all']['quantize_pack_b2_bs64']['frac'], d['roofline_all']['quantize_pack_b2_whole']['frac'], d['clocks'])
except Exception as e: print('ERR', e)
PY
done
