"""Latency of the public calls on ONE 60 s clip (BASELINE cfg 1): NumPy in -> NumPy out, median of 20 calls.

    python tools/latency_cfg1.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, librosa_b200 as lb
y = (0.1*np.random.default_rng(0).standard_normal(1323000)).astype(np.float32)
for name, fn in [("stft", lambda: lb.stft(y)), ("mel", lambda: lb.feature.melspectrogram(y=y, sr=22050)), ("mfcc", lambda: lb.feature.mfcc(y=y, sr=22050)),
                 ("centroid", lambda: lb.feature.spectral_centroid(y=y, sr=22050)), ("istft", None)]:
    if name == "istft":
        D = lb.stft(y); fn = lambda: lb.istft(D, length=len(y))
    for _ in range(5): fn()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(name, "median ms", round(1e3*sorted(ts)[10], 3), "min", round(1e3*min(ts), 3))
