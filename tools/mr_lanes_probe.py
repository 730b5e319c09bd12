import json, os, sys
sys.path.insert(0, '.')
import numpy as np, librosa_b200 as lb
ctx = lb.default_context()
yd = ctx.to_device((0.1 * np.random.default_rng(0).standard_normal((512, 160000))).astype(np.float32))
for n_fft, hop in ((960, 240), (1200, 300), (2000, 500), (640, 160)):
    for lanes in ("16", "32"):
        os.environ["B2L_MR_LANES"] = lanes
        fn = lambda: lb.feature.melspectrogram(y=yd, sr=16000, n_fft=n_fft, hop_length=hop, n_mels=80)
        for _ in range(2): fn().free()
        ctx.synchronize()
        e0, e1 = ctx.event(), ctx.event(); e0.record()
        for _ in range(5): fn().free()
        e1.record(); ctx.synchronize()
        print(json.dumps({"n_fft": n_fft, "hop": hop, "lanes": lanes, "ms": round(e0.elapsed_ms(e1) / 5, 3)}), flush=True)
