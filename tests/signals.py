"""Deterministic synthetic signals shared by the golden generator, the tests and bench.py
(mix A = white noise, mix B = chirps/tones + -60 dB noise, mix C = silence + one burst)."""
from __future__ import annotations

import numpy as np


def mix_a(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(1234 + seed)
    return (0.1 * rng.standard_normal(n)).astype(np.float32)


def mix_b(n: int, seed: int, sr: float = 22050.0) -> np.ndarray:
    rng = np.random.default_rng(1234 + seed)
    t = np.arange(n, dtype=np.float64) / sr
    y = np.zeros(n, dtype=np.float64)
    for _ in range(3):
        f0, f1 = rng.uniform(50.0, 0.45 * sr, size=2)
        amp = rng.uniform(0.05, 0.5)
        dur = max(t[-1], 1e-9)
        phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur)
        y += amp * np.sin(phase)
    y += 1e-3 * rng.standard_normal(n)
    return y.astype(np.float32)


def mix_c(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(1234 + seed)
    y = np.zeros(n, dtype=np.float32)
    m = min(1000, n)
    start = int(rng.integers(0, max(1, n - m)))
    y[start : start + m] = (0.1 * rng.standard_normal(m)).astype(np.float32)
    return y


MIXES = {"A": mix_a, "B": mix_b, "C": mix_c}


def make(mix: str, shape, seed: int = 0, sr: float = 22050.0) -> np.ndarray:
    """Array of the given shape (..., n): every leading index is an independent clip (seed + flat index)."""
    shape = tuple(shape)
    n = shape[-1]
    lead = shape[:-1]
    count = int(np.prod(lead)) if lead else 1
    fn = MIXES[mix]
    rows = [fn(n, seed + i, sr) if mix == "B" else fn(n, seed + i) for i in range(count)]
    return np.stack(rows).reshape(shape) if lead else rows[0]
