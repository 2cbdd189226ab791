set -u
mkdir -p gpurun_out
T="timeout -k 5"
S=$SECONDS; $T 900 python bench.py > gpurun_out/v1_bench.json 2> gpurun_out/v1_bench.err; echo "bench rc=$? wall=$((SECONDS-S))s"
S=$SECONDS; $T 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v1_ref.json 2> gpurun_out/v1_ref.err; echo "ref rc=$? wall=$((SECONDS-S))s"; cat gpurun_out/v1_ref.json | cut -c1-300
