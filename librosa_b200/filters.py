"""Window and filter-bank construction (host side, float64) — mirror of the hot-path parts of
librosa/filters.py: ``mel`` (:117-251), ``get_window`` (:915-977), ``window_sumsquare`` (:1268-1339).

These constants are tiny and parity-critical, so they are computed on the host with the same float64
expressions as the reference and uploaded once per plan; the GPU only *applies* them.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.signal

from .core.convert import fft_frequencies, hz_to_octs, mel_frequencies
from .util.exceptions import ParameterError
from .util.utils import normalize, pad_center


def get_window(window, Nx: int, *, fftbins: bool = True) -> np.ndarray:
    """Resolve a window specification to an array of length ``Nx`` (librosa/filters.py:961-977)."""
    if callable(window):
        return window(Nx)
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    if isinstance(window, (np.ndarray, list)):
        if len(window) == Nx:
            return np.asarray(window)
        raise ParameterError(f"Window size mismatch: {len(window):d} != {Nx:d}")
    raise ParameterError(f"Invalid window specification: {window!r}")


def mel(*, sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0, fmax=None, htk: bool = False,
        norm="slaney", dtype=np.float32) -> np.ndarray:
    """Triangular mel filter bank, shape ``(n_mels, 1 + n_fft//2)`` (librosa/filters.py:206-251)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    bin_hz = fft_frequencies(sr=sr, n_fft=n_fft)
    edges = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    widths = np.diff(edges)
    offsets = np.subtract.outer(edges, bin_hz)
    for i in range(n_mels):
        up = -offsets[i] / widths[i]
        down = offsets[i + 2] / widths[i + 1]
        weights[i] = np.maximum(0, np.minimum(up, down))
    if isinstance(norm, str):
        if norm != "slaney":
            raise ParameterError(f"Unsupported norm={norm}")
        weights *= (2.0 / (edges[2 : n_mels + 2] - edges[:n_mels]))[:, np.newaxis]
    else:
        weights = normalize(weights, norm=norm, axis=-1)
    if not np.all((edges[:-2] == 0) | (weights.max(axis=1) > 0)):
        warnings.warn(
            "Empty filters detected in mel frequency basis. "
            "Some channels will produce empty responses. "
            "Try increasing your sampling rate (and fmax) or "
            "reducing n_mels.",
            stacklevel=2,
        )
    return weights


def window_sumsquare(*, window, n_frames: int, hop_length: int = 512, win_length=None, n_fft: int = 2048,
                     dtype=np.float32, norm=None) -> np.ndarray:
    """Sum of squared, hop-shifted windows, length ``n_fft + hop*(n_frames-1)``
    (librosa/filters.py:1325-1339; the fill loop of :1258-1265 is vectorised per frame)."""
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = get_window(window, win_length)
    win_sq = normalize(win_sq, norm=norm) ** 2
    win_sq = pad_center(win_sq, size=n_fft)
    for i in range(n_frames):
        s = i * hop_length
        x[s : min(n, s + n_fft)] += win_sq[: max(0, min(n_fft, n - s))]
    return x


def chroma(*, sr: float, n_fft: int, n_chroma: int = 12, tuning: float = 0.0, ctroct: float = 5.0, octwidth=2,
           norm=2, base_c: bool = True, dtype=np.float32) -> np.ndarray:
    """Chroma filter bank ``(n_chroma, 1 + n_fft/2)`` projecting FFT bins onto pitch classes; host-side
    constant with the values of ``librosa.filters.chroma`` (filters.py:254-392): Gaussian bumps around each
    bin's pitch class, column-normalised, optionally weighted by a Gaussian over octaves."""
    # position of every FFT bin (DC excluded) on the chroma axis, in bins
    bin_hz = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    pos = n_chroma * hz_to_octs(bin_hz, tuning=tuning, bins_per_octave=n_chroma)
    # the 0 Hz bin gets a made-up position 1.5 octaves below bin 1
    pos = np.concatenate(([pos[0] - 1.5 * n_chroma], pos))
    width = np.concatenate((np.maximum(pos[1:] - pos[:-1], 1.0), [1]))
    half = np.round(float(n_chroma) / 2)
    # signed distance of every bin to every pitch class, wrapped to [-n_chroma/2, n_chroma/2)
    dist = np.subtract.outer(pos, np.arange(0, n_chroma, dtype="d")).T
    dist = np.remainder(dist + half + 10 * n_chroma, n_chroma) - half
    wts = np.exp(-0.5 * (2 * dist / np.tile(width, (n_chroma, 1))) ** 2)
    wts = normalize(wts, norm=norm, axis=0)
    if octwidth is not None:
        wts *= np.exp(-0.5 * (((pos / n_chroma - ctroct) / octwidth) ** 2))[np.newaxis, :]
    if base_c:
        wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=dtype)
