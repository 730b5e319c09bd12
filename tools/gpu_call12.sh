set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q > gpurun_out/c12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c12_tests.log
timeout 300 python - > gpurun_out/c12_resample.log 2>&1 <<'P'
import numpy as np, librosa_b200 as lb
ctx = lb.default_context()
y = (0.1 * np.random.default_rng(0).standard_normal((1024, 220500))).astype(np.float32)
yd = ctx.to_device(y)
for (a, b) in ((22050, 16000), (44100, 16000), (22050, 44100)):
    for _ in range(3): lb.resample(yd, orig_sr=a, target_sr=b, res_type="polyphase").free()
    ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event(); e0.record()
    for _ in range(10): lb.resample(yd, orig_sr=a, target_sr=b, res_type="polyphase").free()
    e1.record(); ctx.synchronize()
    print("resample polyphase", a, "->", b, "1024 x 10 s:", round(e0.elapsed_ms(e1) / 10, 3), "ms")
bad = y[:2].copy(); bad[1, 5] = np.nan
try:
    lb.resample(bad, orig_sr=22050, target_sr=16000, res_type="polyphase"); print("no error?!")
except lb.ParameterError as e: print("ParameterError:", e)
P
tail -n 4 gpurun_out/c12_tests.log; cat gpurun_out/c12_resample.log
