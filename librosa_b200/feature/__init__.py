"""``librosa.feature`` names of the FFT time-frequency path."""
from .spectral import melspectrogram, mfcc

__all__ = ["melspectrogram", "mfcc"]
