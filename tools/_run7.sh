set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/r2_t7_features.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_features.log
rm -f gpurun_out/r2_ab7.jsonl
L=$PWD/librosa_b200/csrc
for rep in 1 2; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg4 >> gpurun_out/r2_ab7.jsonl 2>> gpurun_out/r2_ab7.err
B2L_LIB_PATH=$L/libb2l_p1.so timeout 300 python tools/ab_time.py --tag pvec1 cfg2 cfg4 >> gpurun_out/r2_ab7.jsonl 2>> gpurun_out/r2_ab7.err
B2L_LIB_PATH=$L/libb2l_p2.so timeout 300 python tools/ab_time.py --tag pvec2 cfg2 cfg4 >> gpurun_out/r2_ab7.jsonl 2>> gpurun_out/r2_ab7.err
B2L_LIB_PATH=$L/libb2l_nd.so timeout 300 python tools/ab_time.py --tag nodefer cfg2 cfg4 >> gpurun_out/r2_ab7.jsonl 2>> gpurun_out/r2_ab7.err
B2L_MEL_LPT=0 timeout 300 python tools/ab_time.py --tag nolpt cfg2 cfg4 >> gpurun_out/r2_ab7.jsonl 2>> gpurun_out/r2_ab7.err
done
timeout 600 python tools/feature_timing.py 1024 10 2>&1 >/dev/null | grep -i "rms\|zero" > gpurun_out/r2_feature_timing_zcr2.log
tail -n 4 gpurun_out/r2_t7_features.log; cat gpurun_out/r2_ab7.jsonl | cut -c1-120; cat gpurun_out/r2_feature_timing_zcr2.log | cut -c1-160
