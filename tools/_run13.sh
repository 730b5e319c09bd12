set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=$PWD/librosa_b200/csrc
B2L_LIB_PATH=$L/libb2l_pp.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -q -x > gpurun_out/r2_t13_pp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t13_pp.log
rm -f gpurun_out/r2_ab13.jsonl
for rep in 1 2 3; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg4 stats >> gpurun_out/r2_ab13.jsonl 2>> gpurun_out/r2_ab13.err
B2L_LIB_PATH=$L/libb2l_pp.so timeout 300 python tools/ab_time.py --tag pipe cfg2 cfg4 stats >> gpurun_out/r2_ab13.jsonl 2>> gpurun_out/r2_ab13.err
done
B2L_LIB_PATH=$L/libb2l_pp.so timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck13.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t13_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t13_full.log
cat gpurun_out/r2_ab13.jsonl | cut -c1-110; tail -n 3 gpurun_out/r2_t13_pp.log gpurun_out/r2_t13_full.log; tail -n 2 gpurun_out/r2_racecheck13.log
