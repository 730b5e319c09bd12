set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt
B2L_TMEM=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t1_tm0.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1_tm0.log
B2L_TMEM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t1_tm1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1_tm1.log
rm -f gpurun_out/r2_ab1.jsonl
for tm in 0 1 0 1; do
  B2L_TMEM=$tm timeout 300 python tools/ab_time.py --tag tm$tm cfg2 cfg3 cfg4 cfg5 >> gpurun_out/r2_ab1.jsonl 2>> gpurun_out/r2_ab1.err
done
B2L_TMEM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_tm1 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_tm1.log 2>&1
B2L_TMEM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_tm0 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_tm0.log 2>&1
tail -3 gpurun_out/r2_t1_tm0.log gpurun_out/r2_t1_tm1.log; cat gpurun_out/r2_ab1.jsonl
