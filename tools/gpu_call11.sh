set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/c11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c11_tests.log
timeout 600 python tools/feature_timing.py 1024 10 > gpurun_out/c11_feature_timing.json 2> gpurun_out/c11_feature_timing.log
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np, librosa_b200 as lb
y=(0.1*np.random.default_rng(0).standard_normal((3,9003))).astype(np.float32)
lb.feature.chroma_stft(y=y, sr=22050, tuning=0.0); lb.feature.chroma_stft(y=y, sr=22050, n_fft=1024, hop_length=256, n_chroma=24); lb.feature.chroma_stft(y=y, sr=16000, n_fft=400, hop_length=160, tuning=0.1)
print('sanitize ok')
" > gpurun_out/c11_sanitize.log 2>&1
tail -n 3 gpurun_out/c11_tests.log; tail -2 gpurun_out/c11_sanitize.log; grep -A3 'chroma' gpurun_out/c11_feature_timing.json | head -12
