set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t16_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t16_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke16.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke16.log
rm -f gpurun_out/r2_ab16.jsonl
for rep in 1 2; do timeout 300 python tools/ab_time.py --tag v11 cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/r2_ab16.jsonl 2>> gpurun_out/r2_ab16.err; done
timeout 900 python bench.py > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench16_ref.json 2> gpurun_out/r2_bench16_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_v11 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_v11.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_stft4096_v3 python tools/prof_run.py cfg3 5 512 > gpurun_out/r2_ncu_stft3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_cfg2_v11.csv python tools/prof_run.py cfg2 6 > /dev/null 2>&1
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/r2_feature_timing_v11.json 2> gpurun_out/r2_feature_timing_v11.log
tail -n 4 gpurun_out/r2_t16_full.log gpurun_out/r2_smoke16.log; cat gpurun_out/r2_ab16.jsonl | cut -c1-110; head -c 300 gpurun_out/r2_bench16.json
