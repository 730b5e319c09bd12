set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
ALT=$PWD/librosa_b200/csrc/libb2l_alt.so
B2L_LIB_PATH=$ALT timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/c3_tests_alt.log 2>&1; echo "rc=$?" >> gpurun_out/c3_tests_alt.log
rm -f gpurun_out/c3_ab.jsonl
for rep in 1 2; do
  timeout 300 python tools/ab_time.py --tag main cfg2 cfg4 >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3_ab.err
  B2L_LIB_PATH=$ALT timeout 300 python tools/ab_time.py --tag alt_seg cfg2 cfg4 >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3_ab.err
  B2L_LIB_PATH=$ALT B2L_MEL_SEG=0 timeout 300 python tools/ab_time.py --tag alt_rows_packed cfg2 cfg4 >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3_ab.err
done
B2L_LIB_PATH=$ALT timeout 300 python tools/time_nonpow2.py > gpurun_out/c3_nonpow2.jsonl 2> gpurun_out/c3_nonpow2.err
B2L_LIB_PATH=$ALT timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -f -o gpurun_out/c3_mel_seg python tools/prof_run.py cfg2 5 > gpurun_out/c3_ncu_mel.log 2>&1
B2L_LIB_PATH=$ALT timeout 400 ncu --set full --clock-control none --import-source on -k regex:mr_kernel -s 3 -c 1 -f -o gpurun_out/c3_mr python -c "
import numpy as np, librosa_b200 as lb
ctx=lb.default_context()
yd=ctx.to_device((0.1*np.random.default_rng(0).standard_normal((1024,160000))).astype(np.float32))
for _ in range(5): lb.feature.melspectrogram(y=yd,sr=16000,n_fft=400,hop_length=160,n_mels=80).free()
ctx.synchronize()
" > gpurun_out/c3_ncu_mr.log 2>&1
# accuracy of the two non-power-of-two front ends on the noise-floor case (mix B + lifter)
B2L_LIB_PATH=$ALT timeout 120 python - > gpurun_out/c3_lifterB.log 2>&1 <<'P'
import os, sys
sys.path.insert(0, 'tests')
import numpy as np, signals, librosa_b200 as lb
from oracle import ref_np as O
y = signals.make("B", (2, 8000), seed=23, sr=16000)
kw = dict(sr=16000, n_mfcc=13, n_fft=400, hop_length=160, n_mels=80, lifter=22)
ref = O.mfcc(y=y, **kw)
for mr in ("1", "0"):
    os.environ["B2L_MR"] = mr
    got = lb.feature.mfcc(y=y, **kw)
    print("B2L_MR", mr, "max abs err", float(np.abs(got - ref).max()))
kw2 = dict(kw); kw2["n_fft"] = 512
print("n_fft 512 max abs err", float(np.abs(lb.feature.mfcc(y=y, **kw2) - O.mfcc(y=y, **kw2)).max()))
P
tail -n 3 gpurun_out/c3_tests_alt.log; cut -c1-120 gpurun_out/c3_ab.jsonl; cat gpurun_out/c3_nonpow2.jsonl gpurun_out/c3_lifterB.log
