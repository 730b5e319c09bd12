set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=$PWD/librosa_b200/csrc
B2L_LIB_PATH=$L/libb2l_af.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -q -x > gpurun_out/r2_t15_af.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t15_af.log
rm -f gpurun_out/r2_ab15.jsonl
for rep in 1 2 3; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg3 cfg5 stats >> gpurun_out/r2_ab15.jsonl 2>> gpurun_out/r2_ab15.err
B2L_LIB_PATH=$L/libb2l_af.so timeout 300 python tools/ab_time.py --tag af cfg2 cfg3 cfg5 stats >> gpurun_out/r2_ab15.jsonl 2>> gpurun_out/r2_ab15.err
done
B2L_LIB_PATH=$L/libb2l_af.so timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_memcheck15.log 2>&1
cat gpurun_out/r2_ab15.jsonl | cut -c1-110; tail -n 3 gpurun_out/r2_t15_af.log; tail -n 2 gpurun_out/r2_memcheck15.log
