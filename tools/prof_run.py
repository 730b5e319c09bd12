"""Tiny driver for ncu captures: runs one workload device-resident a few times.
    python tools/prof_run.py cfg2 6"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import librosa_b200 as lb

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 3:
    w["clips"] = int(sys.argv[3])
ctx = lb.default_context()
dev = ctx.to_device(bench.make_batch(w, 0))
kw, op, sr = w["kw"], w["op"], w["sr"]
for _ in range(reps):
    if op == "mel":
        lb.feature.melspectrogram(y=dev, sr=sr, **kw).free()
    elif op == "stft":
        lb.stft(dev, **kw).free()
    elif op == "mfcc":
        lb.feature.mfcc(y=dev, sr=sr, **kw).free()
    elif op == "centroid":
        lb.feature.spectral_centroid(y=dev, sr=sr, **kw).free()
    else:
        D = lb.stft(dev, **kw); lb.istft(D, hop_length=kw["hop_length"], length=w["n"]).free(); D.free()
ctx.synchronize()
print("done", name, reps)
