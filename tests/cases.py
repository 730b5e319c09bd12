"""The parity case table: one entry per (function, parameters, input) combination.

Used three ways:
  * tools/make_golden.py runs every case through the UNMODIFIED reference (build container only) and
    stores the outputs in tests/golden/hotpath_v1.npz;
  * tests/test_oracle_golden.py (CPU) checks the oracle restatement against those stored outputs;
  * tests/test_gpu_parity.py (-m gpu) checks the CUDA path against the oracle *and* the stored outputs.
"""
from __future__ import annotations

CASES = [
    # ---- stft / istft round: (name, op, input spec, kwargs)
    dict(name="stft_2048_512_A", op="stft", mix="A", shape=(9000,), kw=dict(n_fft=2048, hop_length=512)),
    dict(name="stft_1024_256_reflect_B", op="stft", mix="B", shape=(5000,), kw=dict(n_fft=1024, hop_length=256, pad_mode="reflect")),
    dict(name="stft_512_nocenter_A", op="stft", mix="A", shape=(4000,), kw=dict(n_fft=512, center=False)),
    dict(name="stft_2048_short_warns", op="stft", mix="A", shape=(1000,), kw=dict(n_fft=2048, hop_length=512)),
    dict(name="stft_256_edge_A", op="stft", mix="A", shape=(3000,), kw=dict(n_fft=256, hop_length=64, pad_mode="edge")),
    dict(name="stft_1024_symmetric_oddhop_A", op="stft", mix="A", shape=(7000,), kw=dict(n_fft=1024, hop_length=300, pad_mode="symmetric")),
    dict(name="stft_256_linear_ramp_A", op="stft", mix="A", shape=(6000,), kw=dict(n_fft=256, hop_length=64, pad_mode="linear_ramp")),
    dict(name="stft_4096_1024_A", op="stft", mix="A", shape=(12000,), kw=dict(n_fft=4096, hop_length=1024)),
    dict(name="stft_8192_2048_A", op="stft", mix="A", shape=(30000,), kw=dict(n_fft=8192, hop_length=2048)),
    dict(name="stft_64_16_B", op="stft", mix="B", shape=(2000,), kw=dict(n_fft=64, hop_length=16)),
    dict(name="stft_16_4_A", op="stft", mix="A", shape=(500,), kw=dict(n_fft=16, hop_length=4)),
    dict(name="stft_8_2_A", op="stft", mix="A", shape=(200,), kw=dict(n_fft=8, hop_length=2, center=False)),
    dict(name="stft_1024_winlen600_hamming_A", op="stft", mix="A", shape=(5000,), kw=dict(n_fft=1024, win_length=600, window="hamming")),
    dict(name="stft_512_stereo_A", op="stft", mix="A", shape=(2, 3000), kw=dict(n_fft=512, hop_length=128)),
    dict(name="stft_2048_C_burst", op="stft", mix="C", shape=(9000,), kw=dict(n_fft=2048, hop_length=512)),
    # n_fft that is not a power of two (chirp-z kernel): the sizes the reference's own tests use, the 400/160
    # of speech front ends, tiny and near-maximal sizes
    dict(name="stft_501_nonpow2", op="stft", mix="A", shape=(3000,), kw=dict(n_fft=501, hop_length=128)),
    dict(name="stft_1025_nonpow2", op="stft", mix="A", shape=(5000,), kw=dict(n_fft=1025, hop_length=300)),
    dict(name="stft_1023_reflect_B", op="stft", mix="B", shape=(6000,), kw=dict(n_fft=1023, hop_length=256, pad_mode="reflect")),
    dict(name="stft_400_160_stereo_A", op="stft", mix="A", shape=(2, 8000), kw=dict(n_fft=400, hop_length=160)),
    dict(name="stft_2000_500_nocenter_A", op="stft", mix="A", shape=(9000,), kw=dict(n_fft=2000, hop_length=500, center=False)),
    dict(name="stft_6_2_A", op="stft", mix="A", shape=(100,), kw=dict(n_fft=6, hop_length=2)),
    dict(name="stft_12_5_edge_A", op="stft", mix="A", shape=(300,), kw=dict(n_fft=12, hop_length=5, pad_mode="edge")),
    dict(name="stft_600_winlen400_hamming_A", op="stft", mix="A", shape=(4000,), kw=dict(n_fft=600, win_length=400, window="hamming")),
    # even sizes whose half is 5-smooth: the mixed-radix kernel (mr_kernel.cuh) — every radix (5, 3, 8, 4, 2) and
    # several pass counts, all pad modes, win_length < n_fft, center=False
    dict(name="stft_320_80_reflect_B", op="stft", mix="B", shape=(4000,), kw=dict(n_fft=320, hop_length=80, pad_mode="reflect")),
    dict(name="stft_480_120_symmetric_A", op="stft", mix="A", shape=(2, 5000), kw=dict(n_fft=480, hop_length=120, pad_mode="symmetric")),
    dict(name="stft_486_oddhop_A", op="stft", mix="A", shape=(4001,), kw=dict(n_fft=486, hop_length=97)),
    dict(name="stft_250_nocenter_B", op="stft", mix="B", shape=(3000,), kw=dict(n_fft=250, hop_length=50, center=False)),
    dict(name="stft_1200_300_A", op="stft", mix="A", shape=(9000,), kw=dict(n_fft=1200, hop_length=300)),
    dict(name="stft_96_24_linear_ramp_A", op="stft", mix="A", shape=(1500,), kw=dict(n_fft=96, hop_length=24, pad_mode="linear_ramp")),
    dict(name="stft_1920_480_C", op="stft", mix="C", shape=(9000,), kw=dict(n_fft=1920, hop_length=480)),
    dict(name="stft_800_winlen640_edge_A", op="stft", mix="A", shape=(6000,), kw=dict(n_fft=800, win_length=640, hop_length=160, pad_mode="edge")),
    # beyond the chirp-z range (n_fft > 2047): served by the mixed-radix kernels alone
    dict(name="stft_3000_750_A", op="stft", mix="A", shape=(2, 12000), kw=dict(n_fft=3000, hop_length=750)),
    dict(name="stft_4000_1000_nocenter_B", op="stft", mix="B", shape=(15000,), kw=dict(n_fft=4000, hop_length=1000, center=False)),
    # reference-supported, GPU kernels not built: the CUDA path must refuse loudly (oracle still pinned)
    dict(name="stft_3001_toolarge", op="stft", mix="A", shape=(9000,), kw=dict(n_fft=3001, hop_length=700)),   # beyond the chirp-z range: runs on the FP64 kernels
    # ---- istft: input is the golden stft of the named case
    dict(name="istft_2048_512", op="istft", src="stft_2048_512_A", kw=dict(hop_length=512)),
    dict(name="istft_2048_512_length", op="istft", src="stft_2048_512_A", kw=dict(hop_length=512, length=9000)),
    dict(name="istft_1024_256", op="istft", src="stft_1024_256_reflect_B", kw=dict(hop_length=256, length=5000)),
    dict(name="istft_512_nocenter", op="istft", src="stft_512_nocenter_A", kw=dict(center=False)),
    dict(name="istft_1024_oddhop", op="istft", src="stft_1024_symmetric_oddhop_A", kw=dict(hop_length=300)),
    dict(name="istft_1024_winlen600", op="istft", src="stft_1024_winlen600_hamming_A", kw=dict(win_length=600, window="hamming")),
    dict(name="istft_512_stereo", op="istft", src="stft_512_stereo_A", kw=dict(hop_length=128)),
    dict(name="istft_64_16", op="istft", src="stft_64_16_B", kw=dict(hop_length=16, length=2000)),
    dict(name="istft_1025_nonpow2", op="istft", src="stft_1025_nonpow2", kw=dict(hop_length=300, n_fft=1025)),
    dict(name="istft_501_length", op="istft", src="stft_501_nonpow2", kw=dict(hop_length=128, n_fft=501, length=3000)),
    dict(name="istft_400_160_stereo", op="istft", src="stft_400_160_stereo_A", kw=dict(hop_length=160, length=8000)),
    dict(name="istft_2000_nocenter", op="istft", src="stft_2000_500_nocenter_A", kw=dict(hop_length=500, center=False)),
    dict(name="istft_1200_300", op="istft", src="stft_1200_300_A", kw=dict(hop_length=300, length=9000)),
    dict(name="istft_480_120_stereo", op="istft", src="stft_480_120_symmetric_A", kw=dict(hop_length=120)),
    dict(name="istft_250_nocenter", op="istft", src="stft_250_nocenter_B", kw=dict(hop_length=50, center=False)),
    dict(name="istft_486_oddhop", op="istft", src="stft_486_oddhop_A", kw=dict(hop_length=97, length=4001)),
    dict(name="istft_800_winlen640", op="istft", src="stft_800_winlen640_edge_A", kw=dict(hop_length=160, win_length=640, length=6000)),
    dict(name="istft_3000_750", op="istft", src="stft_3000_750_A", kw=dict(hop_length=750, length=12000)),
    dict(name="istft_6_2", op="istft", src="stft_6_2_A", kw=dict(hop_length=2)),
    dict(name="istft_8192_2048", op="istft", src="stft_8192_2048_A", kw=dict(hop_length=2048, length=30000)),
    dict(name="istft_4096_1024_short_length", op="istft", src="stft_4096_1024_A", kw=dict(hop_length=1024, length=7000)),
    # ---- melspectrogram
    dict(name="mel_22050_2048_A", op="mel", mix="A", shape=(9000,), kw=dict(sr=22050, n_fft=2048, hop_length=512)),
    dict(name="mel_22050_2048_B", op="mel", mix="B", shape=(9000,), kw=dict(sr=22050, n_fft=2048, hop_length=512)),
    dict(name="mel_22050_2048_C", op="mel", mix="C", shape=(9000,), kw=dict(sr=22050, n_fft=2048, hop_length=512)),
    dict(name="mel_16000_1024_stereo_A", op="mel", mix="A", shape=(2, 6000), kw=dict(sr=16000, n_fft=1024, hop_length=256)),
    dict(name="mel_44100_4096_A", op="mel", mix="A", shape=(14000,), kw=dict(sr=44100, n_fft=4096, hop_length=1024)),
    dict(name="mel_48000_8192_A", op="mel", mix="A", shape=(40000,), kw=dict(sr=48000, n_fft=8192, hop_length=2048)),
    dict(name="mel_htk_40_power1_A", op="mel", mix="A", shape=(6000,), kw=dict(sr=22050, n_fft=1024, hop_length=256, n_mels=40, htk=True, power=1.0)),
    dict(name="mel_norm1_fminfmax_B", op="mel", mix="B", shape=(6000,), kw=dict(sr=22050, n_fft=2048, hop_length=512, n_mels=64, fmin=300.0, fmax=8000.0, norm=1)),
    dict(name="mel_16000_400_80_B", op="mel", mix="B", shape=(2, 8000), kw=dict(sr=16000, n_fft=400, hop_length=160, n_mels=80)),
    dict(name="mel_22050_1025_A", op="mel", mix="A", shape=(6000,), kw=dict(sr=22050, n_fft=1025, hop_length=256, n_mels=40)),
    dict(name="mel_48000_960_64_B", op="mel", mix="B", shape=(2, 9000), kw=dict(sr=48000, n_fft=960, hop_length=480, n_mels=64)),
    dict(name="mel_16000_800_power1_A", op="mel", mix="A", shape=(7000,), kw=dict(sr=16000, n_fft=800, hop_length=160, n_mels=80, power=1.0)),
    dict(name="mel_16000_400_128_C", op="mel", mix="C", shape=(3, 6000), kw=dict(sr=16000, n_fft=400, hop_length=160, n_mels=128)),
    dict(name="mel_44100_2400_64_B", op="mel", mix="B", shape=(14000,), kw=dict(sr=44100, n_fft=2400, hop_length=600, n_mels=64)),
    dict(name="mel_power3_A", op="mel", mix="A", shape=(4000,), kw=dict(sr=22050, n_fft=512, hop_length=128, n_mels=32, power=3.0)),
    # ---- mfcc
    dict(name="mfcc_16000_1024_A", op="mfcc", mix="A", shape=(8000,), kw=dict(sr=16000, n_mfcc=40, n_fft=1024, hop_length=256)),
    dict(name="mfcc_16000_1024_B", op="mfcc", mix="B", shape=(8000,), kw=dict(sr=16000, n_mfcc=40, n_fft=1024, hop_length=256)),
    dict(name="mfcc_22050_2048_C_clamped", op="mfcc", mix="C", shape=(9000,), kw=dict(sr=22050, n_mfcc=20)),
    dict(name="mfcc_stereo_perchannel_max", op="mfcc", mix="C", shape=(2, 9000), kw=dict(sr=22050, n_mfcc=13)),
    dict(name="mfcc_16000_400_C", op="mfcc", mix="C", shape=(2, 8000), kw=dict(sr=16000, n_mfcc=13, n_fft=400, hop_length=160, n_mels=40)),
    dict(name="mfcc_16000_320_A", op="mfcc", mix="A", shape=(7000,), kw=dict(sr=16000, n_mfcc=20, n_fft=320, hop_length=160, n_mels=40)),
    dict(name="mfcc_16000_400_lifter_A", op="mfcc", mix="A", shape=(2, 8000), kw=dict(sr=16000, n_mfcc=13, n_fft=400, hop_length=160, n_mels=80, lifter=22)),
    dict(name="mfcc_lifter22_dct3", op="mfcc", mix="A", shape=(6000,), kw=dict(sr=22050, n_mfcc=13, lifter=22, dct_type=3)),
    dict(name="mfcc_dct1_nonorm", op="mfcc", mix="A", shape=(6000,), kw=dict(sr=22050, n_mfcc=13, dct_type=1, norm=None)),
]

BY_NAME = {c["name"]: c for c in CASES}
