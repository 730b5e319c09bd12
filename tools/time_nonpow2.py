"""Device-resident timing of the chirp-z path: melspectrogram n_fft=400 / hop=160 / 80 mels, 1024 x 10 s @ 16 kHz."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, librosa_b200 as lb
ctx = lb.default_context()
y = (0.1 * np.random.default_rng(0).standard_normal((1024, 160000))).astype(np.float32)
yd = ctx.to_device(y)
for name, fn in [("mel 400/160/80", lambda: lb.feature.melspectrogram(y=yd, sr=16000, n_fft=400, hop_length=160, n_mels=80)),
                 ("stft 400/160", lambda: lb.stft(yd, n_fft=400, hop_length=160)),
                 ("mel 512/160/80", lambda: lb.feature.melspectrogram(y=yd, sr=16000, n_fft=512, hop_length=160, n_mels=80))]:
    for _ in range(3): fn().free()
    ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(10): fn().free()
    e1.record(); ctx.synchronize()
    print(name, round(e0.elapsed_ms(e1) / 10, 3), "ms per 1024 x 10 s (", 1024 * 1001, "frames )")
