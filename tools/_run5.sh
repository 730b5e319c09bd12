set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t5_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5_full.log
rm -f gpurun_out/r2_ab5.jsonl
for a in 1 2; do B2L_INV2_AHEAD=$a timeout 300 python tools/ab_time.py --tag ahead$a cfg5 >> gpurun_out/r2_ab5.jsonl 2>> gpurun_out/r2_ab5.err; done
for f in 1 2; do B2L_DCT_FPL=$f timeout 300 python tools/ab_time.py --tag fpl$f cfg4 >> gpurun_out/r2_ab5.jsonl 2>> gpurun_out/r2_ab5.err; done
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/r2_feature_timing.json 2> gpurun_out/r2_feature_timing.log
B2L_TD_BLOCK=0 timeout 600 python tools/feature_timing.py 1024 5 2>&1 | grep -i "rms\|zero" > gpurun_out/r2_feature_timing_tdold.log
# launch lists (shares of a step) and one full capture of the cfg-3 stft kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2_launches_cfg2.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cfg4.csv python tools/prof_run.py cfg4 4 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cfg5.csv python tools/prof_run.py cfg5 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_stft4096 python tools/prof_run.py cfg3 5 512 > gpurun_out/r2_ncu_stft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv2_kernel -s 3 -c 1 -f -o gpurun_out/r2_inv2_v2 python tools/prof_run.py cfg5 5 > gpurun_out/r2_ncu_inv2b.log 2>&1
tail -n 6 gpurun_out/r2_t5_full.log; cat gpurun_out/r2_ab5.jsonl | cut -c1-150; cat gpurun_out/r2_feature_timing.log gpurun_out/r2_feature_timing_tdold.log | cut -c1-200
