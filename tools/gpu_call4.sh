set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c4_tests.log
timeout 300 python tools/time_nonpow2.py > gpurun_out/c4_nonpow2.jsonl 2> gpurun_out/c4_nonpow2.err
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np, librosa_b200 as lb
y=(0.1*np.random.default_rng(0).standard_normal((3,9003))).astype(np.float32)
for kw in (dict(n_fft=400,hop_length=160),dict(n_fft=1200,hop_length=301),dict(n_fft=96,hop_length=24,center=False),dict(n_fft=3000,hop_length=750)):
    D=lb.stft(y,**kw); lb.istft(D,hop_length=kw['hop_length'],n_fft=kw['n_fft'],center=kw.get('center',True))
print('sanitize ok')
" > gpurun_out/c4_sanitize.log 2>&1
timeout 900 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -n 3 gpurun_out/c4_tests.log; cat gpurun_out/c4_nonpow2.jsonl; tail -3 gpurun_out/c4_sanitize.log; head -c 600 gpurun_out/c4_bench.json
