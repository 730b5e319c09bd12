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


def mix_t(n: int, seed: int, sr: float = 22050.0) -> np.ndarray:
    """Tonal mix for the tuning / chroma cases: a handful of harmonic notes, all detuned by the same fraction
    of a semitone (0.27 + 0.1 * (seed % 3)), over -50 dB noise — the residual histogram of
    ``estimate_tuning`` then has one clear peak instead of the near-ties noise produces."""
    rng = np.random.default_rng(4321 + seed)
    t = np.arange(n, dtype=np.float64) / sr
    detune = 0.27 + 0.1 * (seed % 3)
    y = np.zeros(n, dtype=np.float64)
    for midi in rng.choice(np.arange(50, 84), size=6, replace=False):
        f0 = 440.0 * 2.0 ** ((midi - 69 + detune) / 12.0)
        for h in range(1, 5):
            if h * f0 < 0.45 * sr:
                y += (0.2 / h) * np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi))
    y += 3e-3 * rng.standard_normal(n)
    return (0.3 * y).astype(np.float32)


MIXES = {"A": mix_a, "B": mix_b, "C": mix_c, "T": mix_t}


def make(mix: str, shape, seed: int = 0, sr: float = 22050.0) -> np.ndarray:
    """Array of the given shape (..., n): every leading index is an independent clip (seed + flat index)."""
    shape = tuple(shape)
    n = shape[-1]
    lead = shape[:-1]
    count = int(np.prod(lead)) if lead else 1
    fn = MIXES[mix]
    rows = [fn(n, seed + i, sr) if mix in ("B", "T") else fn(n, seed + i) for i in range(count)]
    return np.stack(rows).reshape(shape) if lead else rows[0]
