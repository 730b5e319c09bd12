set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
./tools/micro/contrast_test > gpurun_out/contrast_test.log 2>&1
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/c10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c10_tests.log
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/c10_feature_timing.json 2> gpurun_out/c10_feature_timing.log
cat gpurun_out/contrast_test.log; tail -n 3 gpurun_out/c10_tests.log; grep -A3 '"spectral_contrast"' gpurun_out/c10_feature_timing.json
