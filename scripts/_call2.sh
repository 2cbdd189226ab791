set -u
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python -m pytest tests/test_gpu_quantizer.py tests/test_nf_quantizer.py tests/test_bbint_quantizer.py -x -q -m gpu > gpurun_out/u3_pytest_quant.log 2>&1; echo "pytest quant rc=$?"
tail -3 gpurun_out/u3_pytest_quant.log
CB_LIBRARY=$PWD/ee274_convexcaldera_llm_quantization_b200/libcaldera_b200_measure.so $T 200 python scripts/probe_quant_stream.py > gpurun_out/u3_quant_stream.log 2>&1; echo "probe quant rc=$?"
cat gpurun_out/u3_quant_stream.log | cut -c1-160
