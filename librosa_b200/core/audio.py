"""The I/O edge of the hot path: ``librosa.stream``'s block semantics for in-memory audio.

``librosa.stream`` (librosa/core/audio.py:223-500) reads a sound file in blocks that line up with STFT frames:
a block holds ``block_length`` frames, i.e. ``(block_length - 1) * hop_length + frame_length`` samples, and
consecutive blocks advance by ``block_length * hop_length`` samples (they overlap by ``frame_length - hop_length``),
so that ``stft(block, center=False)`` over the blocks concatenates to ``stft(y, center=False)`` of the whole signal.
Decoding audio files and resampling (soundfile / soxr) are outside the path and outside this repository; this
module provides the same block walk over an array that is already in memory (or a ``DeviceArray``), which is
what a decoder thread would feed.
"""
from __future__ import annotations

from typing import Iterator, Optional

import numpy as np

from .. import _native as nat
from ..util.exceptions import ParameterError
from ..util.utils import is_positive_int


def stream(y, *, block_length: int, frame_length: int, hop_length: int, mono: bool = True, offset: float = 0.0,
           duration: Optional[float] = None, fill_value: Optional[float] = None, sr: Optional[float] = None,
           dtype=np.float32) -> Iterator[np.ndarray]:
    """Yield blocks of ``block_length`` frames of an in-memory signal ``y`` (shape ``(n,)`` or ``(channels, n)``).

    Same arguments and block geometry as ``librosa.stream`` (audio.py:385-500), with the file replaced by an
    array: ``offset`` / ``duration`` are in seconds when ``sr`` is given and in samples otherwise; the last block
    is shorter than the others unless ``fill_value`` is given.  A path (``str`` / ``os.PathLike``) is refused:
    audio decoding is not part of the GPU path and there is no CPU fallback."""
    if not is_positive_int(block_length):
        raise ParameterError(f"block_length={block_length} must be a positive integer")
    if not is_positive_int(frame_length):
        raise ParameterError(f"frame_length={frame_length} must be a positive integer")
    if not is_positive_int(hop_length):
        raise ParameterError(f"hop_length={hop_length} must be a positive integer")
    if sr is not None and not (np.isfinite(sr) and sr > 0):
        raise ParameterError(f"sr={sr} must be a positive number")
    if isinstance(y, (str, bytes)) or hasattr(y, "__fspath__"):
        raise nat.UnsupportedOnGPU("stream: decoding audio files (soundfile / soxr) is outside the GPU path; pass the "
                                   "decoded samples as an array")
    if isinstance(y, nat.DeviceArray):
        raise nat.UnsupportedOnGPU("stream: slice the DeviceArray's host source instead (blocks are host views)")
    y = np.asarray(y)
    if y.ndim not in (1, 2):
        raise ParameterError(f"stream expects a (n,) or (channels, n) array, got shape {y.shape}")
    if mono and y.ndim == 2:
        y = np.mean(y, axis=0)
    scale = float(sr) if sr is not None else 1.0
    n = y.shape[-1]
    start = int(offset * scale) if offset >= 0 else max(0, n - int(abs(offset) * scale))
    stop = n if duration is None else min(n, start + int(duration * scale))
    size = (block_length - 1) * hop_length + frame_length
    advance = block_length * hop_length
    pos = start
    while pos < stop:
        block = y[..., pos: min(pos + size, stop)]
        if block.shape[-1] < size and fill_value is not None:
            pad = [(0, 0)] * (block.ndim - 1) + [(0, size - block.shape[-1])]
            block = np.pad(block, pad, mode="constant", constant_values=fill_value)
        yield np.ascontiguousarray(block, dtype=dtype)
        if pos + size >= stop:
            break
        pos += advance
