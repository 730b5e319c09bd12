set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t11_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t11_full.log
rm -f gpurun_out/r2_ab11.jsonl
for rep in 1 2; do
timeout 300 python tools/ab_time.py --tag base cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/r2_ab11.jsonl 2>> gpurun_out/r2_ab11.err
B2L_MEL_LPT=1 timeout 300 python tools/ab_time.py --tag lpt cfg2 cfg4 >> gpurun_out/r2_ab11.jsonl 2>> gpurun_out/r2_ab11.err
done
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck11.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_v10 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_v10.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_cfg2_v10.csv python tools/prof_run.py cfg2 6 > /dev/null 2>&1
tail -n 4 gpurun_out/r2_t11_full.log; cat gpurun_out/r2_ab11.jsonl | cut -c1-110; tail -n 2 gpurun_out/r2_racecheck11.log; head -c 400 gpurun_out/r2_bench11.json
