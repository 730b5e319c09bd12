set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t9_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t9_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench9_ref.json 2> gpurun_out/r2_bench9_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
timeout 300 python tools/ab_time.py --tag final cfg2 cfg3 cfg4 cfg5 stats > gpurun_out/r2_ab9.jsonl 2> gpurun_out/r2_ab9.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_mel_v9 python tools/prof_run.py cfg2 5 > gpurun_out/r2_ncu_mel_v9.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_cfg2_v9.csv python tools/prof_run.py cfg2 6 > /dev/null 2>&1
tail -n 5 gpurun_out/r2_t9_full.log gpurun_out/r2_smoke.log; cat gpurun_out/r2_ab9.jsonl | cut -c1-120; head -c 600 gpurun_out/r2_bench9.json
