set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
B2L_MEL2=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_grids.py -m gpu -q -x -k "mel or Mel or chirpz or multi" > gpurun_out/r2_t8_mel2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8_mel2.log
rm -f gpurun_out/r2_ab8.jsonl
for rep in 1 2; do
timeout 300 python tools/ab_time.py --tag base cfg2 >> gpurun_out/r2_ab8.jsonl 2>> gpurun_out/r2_ab8.err
B2L_MEL2=1 timeout 300 python tools/ab_time.py --tag mel2 cfg2 >> gpurun_out/r2_ab8.jsonl 2>> gpurun_out/r2_ab8.err
done
B2L_MEL2=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_memcheck_mel2.log 2>&1
B2L_MEL2=1 timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck_mel2.log 2>&1
B2L_MEL2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel2_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel2 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel2.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t8_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8_full.log
tail -n 12 gpurun_out/r2_t8_mel2.log | cut -c1-200; cat gpurun_out/r2_ab8.jsonl | cut -c1-120; tail -n 3 gpurun_out/r2_memcheck_mel2.log gpurun_out/r2_racecheck_mel2.log; tail -n 3 gpurun_out/r2_t8_full.log
