set -u
mkdir -p gpurun_out
for cfg in "7 23" "8 23" "6 23"; do set -- $cfg
  timeout -k 5 400 python bench.py --slots $1 --batch $2 --no-cpu --no-model --no-ref-cuda --no-parity > gpurun_out/u8_bench_$1x$2.json 2> gpurun_out/u8_bench_$1x$2.err; echo "bench $1x$2 rc=$?"
  python - gpurun_out/u8_bench_$1x$2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['e2e']['submit_side_step_seconds'], d['clocks'])
except Exception as e: print('ERR', e)
PY
done
