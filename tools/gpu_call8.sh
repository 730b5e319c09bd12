set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c8_tests.log
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/c8_feature_timing.json 2> gpurun_out/c8_feature_timing.log
rm -f gpurun_out/c8_ab.jsonl
B2L_MEL_LPT=1 timeout 300 python tools/ab_time.py --tag lpt cfg4 cfg2 >> gpurun_out/c8_ab.jsonl 2>> gpurun_out/c8_ab.err
timeout 300 python tools/ab_time.py --tag base cfg4 cfg2 >> gpurun_out/c8_ab.jsonl 2>> gpurun_out/c8_ab.err
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np, librosa_b200 as lb
y=(0.1*np.random.default_rng(0).standard_normal((3,9003))).astype(np.float32)
for kw in (dict(sr=22050),dict(sr=22050,n_fft=1024,hop_length=256,quantile=0.25,fmin=100.0),dict(sr=16000,n_fft=512,hop_length=128,n_bands=4),dict(sr=44100,n_fft=4096,hop_length=1024)):
    lb.feature.spectral_contrast(y=y,**kw)
print('sanitize ok')
" > gpurun_out/c8_sanitize.log 2>&1
tail -n 5 gpurun_out/c8_tests.log; cut -c1-110 gpurun_out/c8_ab.jsonl; tail -3 gpurun_out/c8_sanitize.log; grep -A3 contrast gpurun_out/c8_feature_timing.json | head
