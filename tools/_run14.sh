set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t14_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t14_full.log
rm -f gpurun_out/r2_ab14.jsonl
for rep in 1 2; do
timeout 300 python tools/ab_time.py --tag affine cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/r2_ab14.jsonl 2>> gpurun_out/r2_ab14.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2_stft4096_v2 python tools/prof_run.py cfg3 5 512 > gpurun_out/r2_ncu_stft2.log 2>&1
tail -n 4 gpurun_out/r2_t14_full.log; cat gpurun_out/r2_ab14.jsonl | cut -c1-110
