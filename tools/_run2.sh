set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2_full.log
rm -f gpurun_out/r2_ab2.jsonl
for tm in 0 1; do
  B2L_TMEM=$tm timeout 300 python tools/ab_time.py --tag tm$tm cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/r2_ab2.jsonl 2>> gpurun_out/r2_ab2.err
done
timeout 600 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_v8 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_v8.log 2>&1
tail -n 3 gpurun_out/r2_t2_full.log; cat gpurun_out/r2_ab2.jsonl; tail -n 5 gpurun_out/r2_racecheck.log gpurun_out/r2_memcheck.log; head -c 3000 gpurun_out/r2_bench2.json
