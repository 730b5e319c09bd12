"""Device-resident A/B timing of the bench workloads under the current environment (B2L_* switches).

    python tools/ab_time.py [--reps 30] [--tag name] cfg2 cfg3 cfg4 cfg5

One JSON line per workload: ms per step (CUDA events on the launching stream, after 3 warm-up steps),
frames/s and the algorithmic-byte bandwidth.  A development aid; the contract numbers come from bench.py."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import librosa_b200 as lb

ap = argparse.ArgumentParser()
ap.add_argument("workloads", nargs="*", default=["cfg2"])
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--tag", default="")
ap.add_argument("--clips", type=int, default=0)
args = ap.parse_args()
ctx = lb.default_context()
for name in args.workloads:
    w = dict(bench.WORKLOADS[name])
    if args.clips:
        w["clips"] = args.clips
    dev = ctx.to_device(bench.make_batch(w, 0))
    kw, op, sr = w["kw"], w["op"], w["sr"]

    def step():
        if op == "mel":
            lb.feature.melspectrogram(y=dev, sr=sr, **kw).free()
        elif op == "stft":
            lb.stft(dev, **kw).free()
        elif op == "mfcc":
            lb.feature.mfcc(y=dev, sr=sr, **kw).free()
        elif op == "centroid":
            lb.feature.spectral_centroid(y=dev, sr=sr, **kw).free()
        else:
            D = lb.stft(dev, **kw)
            lb.istft(D, hop_length=kw["hop_length"], length=w["n"]).free()
            D.free()

    for _ in range(3):
        step()
    ctx.synchronize()
    best = 1e9
    tot = 0.0
    rounds = 3
    for _ in range(rounds):
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(args.reps):
            step()
        e1.record()
        ms = e0.elapsed_ms(e1) / args.reps
        best = min(best, ms)
        tot += ms
    frames = w["clips"] * bench.n_frames(w["n"], kw["n_fft"], kw["hop_length"])
    alg = bench.algorithmic_bytes_per_step(w)
    print(json.dumps({"tag": args.tag, "workload": name, "ms": round(tot / rounds, 4), "ms_best": round(best, 4),
                      "mframes_s": round(frames / (tot / rounds) / 1e3, 1), "alg_gbs": round(alg / (tot / rounds) / 1e6, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("B2L_")}}), flush=True)
    dev.free()
