set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t3_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3_full.log
rm -f gpurun_out/r2_ab3.jsonl
timeout 300 python tools/ab_time.py --tag base cfg2 cfg3 cfg4 cfg5 stats >> gpurun_out/r2_ab3.jsonl 2>> gpurun_out/r2_ab3.err
B2L_LIB_PATH=$PWD/librosa_b200/csrc/libb2l_alt.so timeout 300 python tools/ab_time.py --tag shfl cfg2 cfg5 stats >> gpurun_out/r2_ab3.jsonl 2>> gpurun_out/r2_ab3.err
B2L_INV2=0 timeout 300 python tools/ab_time.py --tag inv2off cfg5 >> gpurun_out/r2_ab3.jsonl 2>> gpurun_out/r2_ab3.err
B2L_DCT_CONST=0 timeout 300 python tools/ab_time.py --tag dctsmem cfg4 >> gpurun_out/r2_ab3.jsonl 2>> gpurun_out/r2_ab3.err
B2L_LIB_PATH=$PWD/librosa_b200/csrc/libb2l_alt.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_t3_alt.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3_alt.log
timeout 600 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv2_kernel -s 3 -c 1 -f -o gpurun_out/r2_inv2 python tools/prof_run.py cfg5 5 > gpurun_out/r2_ncu_inv2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dct_clamp -s 3 -c 1 -f -o gpurun_out/r2_dct python tools/prof_run.py cfg4 5 > gpurun_out/r2_ncu_dct.log 2>&1
tail -n 5 gpurun_out/r2_t3_full.log gpurun_out/r2_t3_alt.log; cat gpurun_out/r2_ab3.jsonl; head -c 4000 gpurun_out/r2_bench3.json; tail -n 3 gpurun_out/r2_bench3.err
