"""Device-resident timings of the frame-wise features on a cfg-2 shaped batch, next to the
oracle (CPU, one process) on a small sample.  CUDA events around `reps` calls after warm-up.

    python tools/feature_timing.py [clips=1024] [reps=10] > gpurun_out/feature_timing.json
"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import bench
import librosa_b200 as lb
from oracle import ref_np as O

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = dict(bench.WORKLOADS["cfg2"], clips=clips)
host = bench.make_batch(w, 0)
ctx = lb.default_context()
dev = ctx.to_device(host)
sr = 22050
T = 1 + host.shape[-1] // 512
frames = clips * T
mel_db = lb.power_to_db(lb.feature.melspectrogram(y=dev, sr=sr))
mel_pw = lb.feature.melspectrogram(y=dev, sr=sr)
dev128 = ctx.to_device(host[:128])


def free(x):
    for a in (x if isinstance(x, tuple) else (x,)):
        if isinstance(a, lb.DeviceArray):
            a.free()


FEATURES = {
    "spectral_centroid": (lambda: lb.feature.spectral_centroid(y=dev, sr=sr), lambda y: O.spectral_centroid(y=y, sr=sr)),
    "spectral_bandwidth": (lambda: lb.feature.spectral_bandwidth(y=dev, sr=sr), lambda y: O.spectral_bandwidth(y=y, sr=sr)),
    "spectral_rolloff": (lambda: lb.feature.spectral_rolloff(y=dev, sr=sr), lambda y: O.spectral_rolloff(y=y, sr=sr)),
    "spectral_flatness": (lambda: lb.feature.spectral_flatness(y=dev), lambda y: O.spectral_flatness(y=y)),
    "spectral_contrast": (lambda: lb.feature.spectral_contrast(y=dev, sr=sr), lambda y: O.spectral_contrast(y=y, sr=sr)),
    "rms(y)": (lambda: lb.feature.rms(y=dev), lambda y: O.rms(y=y)),
    "zero_crossing_rate": (lambda: lb.feature.zero_crossing_rate(dev), lambda y: O.zero_crossing_rate(y)),
    "chroma_stft(tuning=0)": (lambda: lb.feature.chroma_stft(y=dev, sr=sr, tuning=0.0), lambda y: O.chroma_stft(y=y, sr=sr, tuning=0.0)),
    "chroma_stft(estimated tuning)": (lambda: lb.feature.chroma_stft(y=dev, sr=sr), lambda y: O.chroma_stft(y=y, sr=sr)),
    "onset_strength": (lambda: lb.onset.onset_strength(y=dev, sr=sr), lambda y: O.onset_strength(y=y, sr=sr)),
    "effects.hpss (128 clips)": (lambda: lb.effects.hpss(dev128), None),
    "pcen(mel)": (lambda: lb.pcen(mel_pw, sr=sr), None),
    "amplitude_to_db(mel)": (lambda: lb.amplitude_to_db(mel_pw), None),
}
out = {"clips": clips, "frames": frames, "reps": reps, "features": {}}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name, (gpu, cpu) in FEATURES.items():
        for _ in range(3):
            free(gpu())
        ctx.synchronize()
        e0, e1 = ctx.event(), ctx.event()
        l0 = ctx.launch_count
        e0.record()
        for _ in range(reps):
            free(gpu())
        e1.record()
        ctx.synchronize()
        ms = e0.elapsed_ms(e1) / reps
        nfr = 128 * T if "128 clips" in name else frames
        row = {"gpu_ms": round(ms, 3), "frames": nfr, "gpu_frames_per_s": round(nfr / ms * 1e3),
               "launches_per_call": (ctx.launch_count - l0) / reps}
        if cpu is not None:
            sample = host[:8]
            cpu(sample[:1])
            t0 = time.perf_counter()
            cpu(sample)
            dt = time.perf_counter() - t0
            row["cpu_oracle_frames_per_s_1proc"] = round(sample.shape[0] * T / dt)
        out["features"][name] = row
        print(name, row, file=sys.stderr)
print(json.dumps(out, indent=1))
