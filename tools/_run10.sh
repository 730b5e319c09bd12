set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
rm -f gpurun_out/r2_ab10.jsonl
L=$PWD/librosa_b200/csrc
for rep in 1 2 3; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg4 stats >> gpurun_out/r2_ab10.jsonl 2>> gpurun_out/r2_ab10.err
B2L_LIB_PATH=$L/libb2l_d2.so timeout 300 python tools/ab_time.py --tag defer2 cfg2 cfg4 stats >> gpurun_out/r2_ab10.jsonl 2>> gpurun_out/r2_ab10.err
done
B2L_LIB_PATH=$L/libb2l_d2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_t10_d2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t10_d2.log
cat gpurun_out/r2_ab10.jsonl | cut -c1-110; tail -n 3 gpurun_out/r2_t10_d2.log
