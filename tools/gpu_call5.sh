set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
export B2L_LIB_PATH=$PWD/librosa_b200/csrc/libb2l_alt.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/c5_tests_alt.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests_alt.log
rm -f gpurun_out/c5_nonpow2.jsonl
for lanes in 16 32 16 32; do B2L_SKIP_CZT=1 B2L_MR_LANES=$lanes timeout 300 python tools/time_nonpow2.py >> gpurun_out/c5_nonpow2.jsonl 2>> gpurun_out/c5_nonpow2.err; done
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np, librosa_b200 as lb
y=(0.1*np.random.default_rng(0).standard_normal((3,5003))).astype(np.float32)
for kw in (dict(n_fft=400,hop_length=160),dict(n_fft=96,hop_length=24,center=False),dict(n_fft=12,hop_length=5)):
    lb.stft(y,**kw); lb.feature.melspectrogram(y=y,sr=16000,n_mels=40,**kw); lb.feature.mfcc(y=y,sr=16000,n_mels=40,n_mfcc=13,**kw)
print('sanitize ok')
" > gpurun_out/c5_sanitize.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mr_kernel -s 3 -c 1 -f -o gpurun_out/c5_mr python -c "
import numpy as np, librosa_b200 as lb
ctx=lb.default_context()
yd=ctx.to_device((0.1*np.random.default_rng(0).standard_normal((1024,160000))).astype(np.float32))
for _ in range(5): lb.feature.melspectrogram(y=yd,sr=16000,n_fft=400,hop_length=160,n_mels=80).free()
ctx.synchronize()
" > gpurun_out/c5_ncu_mr.log 2>&1
tail -n 3 gpurun_out/c5_tests_alt.log; cat gpurun_out/c5_nonpow2.jsonl; tail -3 gpurun_out/c5_sanitize.log
