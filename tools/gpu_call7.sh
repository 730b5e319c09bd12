set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c7_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c7_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c7_smoke.log
rm -f gpurun_out/c7_ab.jsonl
timeout 300 python tools/ab_time.py --tag final cfg2 cfg3 cfg4 cfg5 stats speech400 >> gpurun_out/c7_ab.jsonl 2>> gpurun_out/c7_ab.err
timeout 300 python tools/time_nonpow2.py > gpurun_out/c7_nonpow2.jsonl 2> gpurun_out/c7_nonpow2.err
timeout 900 python bench.py > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c7_launches_cfg2.csv python tools/prof_run.py cfg2 6 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c7_launches_speech400.csv python tools/prof_run.py speech400 6 > /dev/null 2>&1
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/c7_feature_timing.json 2> gpurun_out/c7_feature_timing.log
tail -n 3 gpurun_out/c7_tests.log gpurun_out/c7_smoke.log; cut -c1-110 gpurun_out/c7_ab.jsonl; head -c 300 gpurun_out/c7_bench.json
