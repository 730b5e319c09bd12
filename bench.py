#!/usr/bin/env python
"""bench.py — headline benchmark of the librosa FFT time-frequency hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5]

Metric (BASELINE.json): mel-spectrogram frames/sec, n_fft=2048, hop=512, n_mels=128, float32, on
BASELINE.json configs[1] — batch = 1024 clips x 10 s mono @ 22050 Hz per GPU.  One "step" is one pass of
the fused stft -> |.|^2 -> mel kernel over that batch.  Weak scaling: every rank owns a 1024-clip shard,
no collective on the data path; `value` = frames of all ranks / max-over-ranks device time.

One JSON line on stdout (rank 0).  Extra keys beyond the base contract:
  roofline      dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes
  cpu_baseline  the oracle port (oracle/ref_np.py == the reference's algorithm, bit-exact here) timed on
                this box's host cores on a bounded sample
  e2e           same metric through the public drop-in call with HOST (pinned) buffers, H2D + D2H inside
  clocks        nvidia-smi samples taken during the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (clips per GPU, channels, samples, sr, op, kwargs, algorithmic bytes per frame)
    "cfg2": dict(clips=1024, n=220500, sr=22050, op="mel", kw=dict(n_fft=2048, hop_length=512, n_mels=128, power=2.0),
                 desc="batch=1024 clips x10s mono sr=22050 -> melspectrogram n_fft=2048 hop=512 n_mels=128 power=2.0"),
    "cfg3": dict(clips=1024, n=441000, sr=44100, op="stft", kw=dict(n_fft=4096, hop_length=1024),
                 desc="1024 channel-clips (512 stereo) x10s sr=44100 -> stft n_fft=4096 hop=1024 per GPU"),
    "cfg4": dict(clips=512, n=480000, sr=16000, op="mfcc", kw=dict(n_mfcc=40, n_mels=128, n_fft=1024, hop_length=256),
                 desc="512 clips x30s mono sr=16000 -> mfcc n_mfcc=40 n_mels=128 n_fft=1024 hop=256 per GPU"),
    # frame-wise statistics fused with the stft (cfg-2 shapes); one launch yields all six rows
    "stats": dict(clips=1024, n=220500, sr=22050, op="centroid", kw=dict(n_fft=2048, hop_length=512),
                  desc="batch=1024 clips x10s mono sr=22050 -> spectral_centroid n_fft=2048 hop=512 (fused statistics kernel)"),
    "cfg5": dict(clips=256, n=220500, sr=22050, op="roundtrip", kw=dict(n_fft=2048, hop_length=512),
                 desc="256 clips x10s -> stft -> istft n_fft=2048 hop=512 per GPU"),
}
METRIC = "mel-spectrogram frames/sec (n_fft=2048,hop=512,n_mels=128)"


def n_frames(n, n_fft, hop):
    return 1 + n // hop   # center=True, even n_fft


def algorithmic_bytes_per_step(w):
    """Compulsory HBM traffic of one step: every input sample read once + every output element written once."""
    T = n_frames(w["n"], w["kw"]["n_fft"], w["kw"]["hop_length"])
    F = 1 + w["kw"]["n_fft"] // 2
    clips, n = w["clips"], w["n"]
    if w["op"] == "mel":
        return clips * (4 * n + 4 * w["kw"]["n_mels"] * T)
    if w["op"] == "stft":
        return clips * (4 * n + 8 * F * T)
    if w["op"] == "mfcc":
        return clips * (4 * n + 4 * w["kw"]["n_mfcc"] * T)
    if w["op"] == "roundtrip":
        return clips * (4 * n + 8 * F * T) + clips * (8 * F * T + 4 * n)
    if w["op"] == "centroid":
        return clips * (4 * n + 4 * 6 * T)      # six statistics rows per frame
    raise ValueError(w["op"])


def make_batch(w, rank):
    import signals

    base = signals.make("A", (64, w["n"]), seed=1000 * rank)       # mix A: 0.1 * N(0,1), SURVEY §8d
    reps = -(-w["clips"] // 64)
    scale = (1.0 + 0.01 * np.arange(reps, dtype=np.float32))[:, None, None]
    return np.ascontiguousarray((base[None] * scale).reshape(-1, w["n"])[: w["clips"]])


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 5 ms from a background thread while the
    timed region runs (nvidia-smi -lms cannot sample a region that lasts tens of milliseconds)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, device):
        import threading

        self.samples, self.bits, self.max_mhz, self.power = [], 0, None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            self._stop.wait(0.005)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), samples=len(self.samples),
                       reasons=sorted(k for k, bit in self.REASONS.items() if self.bits & bit),
                       power_w_max=max(self.power) if self.power else None)
        return out


# ------------------------------------------------------------------------------------------- CPU port
def _cpu_clip_job(args):
    op, kw, sr, y = args
    from oracle import ref_np as O

    if op == "mel":
        return O.melspectrogram(y=y, sr=sr, **kw).shape[-1]
    if op == "stft":
        return O.stft(y, **kw).shape[-1]
    if op == "mfcc":
        return O.mfcc(y=y, sr=sr, **kw).shape[-1]
    if op == "centroid":
        return O.spectral_centroid(y=y, sr=sr, **kw).shape[-1]
    D = O.stft(y, **kw)
    O.istft(D, hop_length=kw["hop_length"], length=len(y))
    return D.shape[-1]


def _cpu_worker_init():
    # one BLAS / OpenMP thread per worker process: the pool already uses every core
    global _BLAS_LIMIT
    try:
        import threadpoolctl

        _BLAS_LIMIT = threadpoolctl.threadpool_limits(limits=1)
    except Exception:
        _BLAS_LIMIT = None


class CpuPort:
    """The reference's algorithm (oracle port) on the host cores: per-clip loop spread over a process pool,
    one BLAS thread per worker — the fastest of the three ways SURVEY §8d lists to run the reference
    (a single batched call is ~3x slower than a per-clip loop)."""

    def __init__(self, w, sample_clips):
        import multiprocessing as mp

        self.w = w
        self.cores = os.cpu_count() or 1
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        self.batch = make_batch(dict(w, clips=sample_clips), rank=0)
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_worker_init)
        self.jobs = [(w["op"], w["kw"], w["sr"], self.batch[i]) for i in range(sample_clips)]

    def step(self):
        t0 = time.perf_counter()
        frames = sum(self.pool.map(_cpu_clip_job, self.jobs, chunksize=1))
        return frames, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample_clips(w):
    # sized for roughly 10-30 s of CPU work in total over warm-up + timed steps
    return {"mel": 128, "stft": 128, "mfcc": 64, "roundtrip": 64, "centroid": 128}[w["op"]]


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    # size the per-step sample so that warm-up + K timed steps take about a minute on this box
    probe = CpuPort(w, 16)
    probe.step()
    _, s16 = probe.step()
    probe.close()
    budget_s = 60.0 / (args.steps + max(1, args.warmup))
    clips = int(max(8, min(cpu_sample_clips(w) * 2, 16 * budget_s / max(s16, 1e-3))))
    port = CpuPort(w, clips)
    for _ in range(max(1, args.warmup)):
        port.step()
    frames = secs = 0.0
    for _ in range(args.steps):
        f, s = port.step()
        frames += f
        secs += s
    port.close()
    value = frames / secs
    line = {
        "impl": "reference", "metric": METRIC if w["op"] == "mel" else f"{w['op']} frames/sec", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 in / f64 FFT (reference numerics)", "data": "synthetic",
        "config": {"workload": w["desc"], "name": args.workload},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": port.cores, "kind": "port",
                         "sample": f"{len(port.jobs)} clips of the workload per step, per-clip loop over a "
                                   f"{port.cores}-process pool (1 BLAS thread each)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, w, rank, world, local_rank):
    import librosa_b200 as lb

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    # one process per GPU: keep this process (and the pinned buffers it allocates) on the CPUs local to its GPU,
    # as `numactl --cpunodebind` would; released again before the CPU baseline uses every core
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if os.environ.get("B2L_BENCH_NO_NUMA_BIND") else lb.bind_host_to_device(local_rank)
    ctx = lb.default_context(local_rank)
    kw, op, sr = w["kw"], w["op"], w["sr"]
    T = n_frames(w["n"], kw["n_fft"], kw["hop_length"])
    frames_per_step = w["clips"] * T

    host = lb.pinned_empty((w["clips"], w["n"]), np.float32)
    host[...] = make_batch(w, rank)
    dev = ctx.to_device(host)

    def step_resident():
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

    def step_e2e():
        if op == "mel":
            return lb.feature.melspectrogram(y=host, sr=sr, **kw)
        if op == "stft":
            return lb.stft(host, **kw)
        if op == "mfcc":
            return lb.feature.mfcc(y=host, sr=sr, **kw)
        if op == "centroid":
            return lb.feature.spectral_centroid(y=host, sr=sr, **kw)
        return lb.istft(lb.stft(host, **kw), hop_length=kw["hop_length"], length=w["n"])

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
            import torch

            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident: warm-up, then K steps between events (inputs 0.9 GB >> 126 MB L2: no flush needed)
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    ms = e0.elapsed_ms(e1)
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * frames_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the public call with host buffers (H2D + D2H inside the timed region)
    e2e_steps = max(3, min(args.steps, 10))
    out = step_e2e()                      # warm-up: second stream, plans, device and pinned pools
    d2h = int(out.nbytes)
    out2 = step_e2e()                     # a second result while the first is alive: both pinned buffers exist
    del out, out2
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = step_e2e()
    ctx.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * frames_per_step / e2e_s

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (one launch per step for mel / stft)
    peak, peak_src = measured_peak()
    alg_bytes = algorithmic_bytes_per_step(w)
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = profiled_traffic(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": {"mel": "fwd_kernel<10,32,16,MODE_MEL>", "stft": "fwd_kernel<.,.,.,MODE_STFT>",
                           "mfcc": "fwd_kernel<.,.,.,MODE_MEL>+dct_clamp_kernel", "roundtrip": "fwd_kernel+inv_kernel",
                           "centroid": "fwd_kernel<10,32,16,MODE_STATS>"}[op],
                "note": "kernel time == step time (one launch per step, CUDA events on the launching stream)"}

    # ---- CPU baseline on this box (bounded sample)
    cpu = None
    if numa_cpus:
        os.sched_setaffinity(0, all_cpus)
    if world == 1 and not args.no_cpu:
        port = CpuPort(w, cpu_sample_clips(w))
        port.step()
        f = s = 0.0
        t_end = time.perf_counter() + 12.0
        while True:
            fi, si = port.step()
            f += fi
            s += si
            if time.perf_counter() > t_end:
                break
        port.close()
        cpu = {"value": f / s, "unit": "frames/s", "cores": port.cores, "kind": "port",
               "sample": f"{len(port.jobs)} clips of the workload per pass for ~12 s, per-clip loop over a "
                         f"{port.cores}-process pool (1 BLAS thread each)"}

    line = {
        "metric": METRIC if op == "mel" else f"{op} frames/sec", "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": args.workload, "per_gpu_clips": w["clips"],
                   "frames_per_step_per_gpu": frames_per_step, "parallelism": f"clips sharded x{world}, no collective",
                   "host_cpus_bound_to_gpu": len(numa_cpus) if numa_cpus else None,
                   "l2": "inputs (%.0f MB per step) exceed the 126 MB L2; no flush" % (w["clips"] * w["n"] * 4 / 1e6)},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(host.nbytes),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                "path": "librosa_b200 public call on a pinned host ndarray -> ndarray"},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
