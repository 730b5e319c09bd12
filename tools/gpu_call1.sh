set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c1_smoke.log
rm -f gpurun_out/c1_ab.jsonl
for rep in 1 2; do timeout 300 python tools/ab_time.py --tag v12 cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/c1_mel python tools/prof_run.py cfg2 5 > gpurun_out/c1_ncu_mel.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/c1_stft4096 python tools/prof_run.py cfg3 5 512 > gpurun_out/c1_ncu_stft.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:inv2_kernel -s 3 -c 1 -f -o gpurun_out/c1_inv2 python tools/prof_run.py cfg5 5 > gpurun_out/c1_ncu_inv2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dct -s 3 -c 1 -f -o gpurun_out/c1_dct python tools/prof_run.py cfg4 5 > gpurun_out/c1_ncu_dct.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/c1_mfccfwd python tools/prof_run.py cfg4 5 > gpurun_out/c1_ncu_mfccfwd.log 2>&1
tail -n 3 gpurun_out/c1_tests.log gpurun_out/c1_smoke.log; cut -c1-120 gpurun_out/c1_ab.jsonl
