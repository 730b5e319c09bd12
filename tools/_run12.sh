set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
rm -f gpurun_out/r2_ab12.jsonl
L=$PWD/librosa_b200/csrc
for rep in 1 2 3; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg4 stats >> gpurun_out/r2_ab12.jsonl 2>> gpurun_out/r2_ab12.err
B2L_LIB_PATH=$L/libb2l_d3.so timeout 300 python tools/ab_time.py --tag defer3 cfg2 cfg4 stats >> gpurun_out/r2_ab12.jsonl 2>> gpurun_out/r2_ab12.err
done
B2L_LIB_PATH=$L/libb2l_d3.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -q -x > gpurun_out/r2_t12_d3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t12_d3.log
B2L_LIB_PATH=$L/libb2l_d3.so timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck12.log 2>&1
cat gpurun_out/r2_ab12.jsonl | cut -c1-110; tail -n 3 gpurun_out/r2_t12_d3.log; tail -n 2 gpurun_out/r2_racecheck12.log
