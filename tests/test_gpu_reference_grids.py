"""GPU: the reference's own hot-path test grids, restated against the CUDA path with the reference's dtypes.

Sources (under /root/reference/tests): test_core.py:256-292 (stft == rfft of the windowed frames), :308-314
(frame lengths 2^12 .. 2^16), :317-390 (out= / oversize / undersize), :813-828 (stft -> istft reconstruction,
float64 chirp, atol 1e-6), test_multichannel.py:96-112, 266-285, 653-714 (batch == per channel).  The
reference's fixtures are audio files; synthetic signals of the same kind stand in for them.  float64 inputs run
on the FP64 kernels (csrc/f64_kernels.cuh), so the assertions keep the reference's default tolerances
(np.allclose: rtol 1e-5, atol 1e-8) instead of the float32 ones.
"""
import numpy as np
import pytest
import scipy.fft
import scipy.signal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import librosa_b200

    librosa_b200.default_context()
    return librosa_b200


@pytest.fixture(scope="module")
def y_22050():
    """1.5 s of programme-like material at 22 050 Hz, float64 (stands in for the reference's audio fixture)."""
    rng = np.random.default_rng(22050)
    t = np.arange(int(1.5 * 22050)) / 22050.0
    y = 0.4 * np.sin(2 * np.pi * 220 * t) + 0.2 * scipy.signal.chirp(t, 100, t[-1], 6000) + 0.05 * rng.standard_normal(t.size)
    return y.astype(np.float64)


def chirp(sr, duration=2.0, fmin=32.0, fmax=8192.0):
    """librosa.chirp(fmin, fmax, sr, duration): exponential sweep, float64 (librosa/core/audio.py:1432)."""
    t = np.arange(int(np.ceil(duration * sr))) / sr
    return scipy.signal.chirp(t, fmin, duration, fmax, method="logarithmic", phi=-90.0)


# ------------------------------------------------------------------ test_core.py:256-292
@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("n_fft", [256, 501])
@pytest.mark.parametrize("window", ["hann", "ones"])
@pytest.mark.parametrize("hop_length", [None, 128])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_stft_equals_rfft_of_frames(lb, y_22050, n_fft, window, hop_length, center, dtype):
    y = y_22050.astype(dtype)
    D = lb.stft(y, n_fft=n_fft, window=window, hop_length=hop_length, center=center)
    assert D.ndim == 2 and D.shape[0] == n_fft // 2 + 1
    assert D.dtype == (np.complex128 if dtype == np.float64 else np.complex64)
    hop = n_fft // 4 if hop_length is None else hop_length
    assert D.shape[-1] == (1 + len(y) // hop if center else 1 + (len(y) - n_fft) // hop)
    win = lb.filters.get_window(window, n_fft, fftbins=True)
    src = np.pad(y, n_fft // 2, mode="constant") if center else y
    frames = lb.util.frame(src, frame_length=n_fft, hop_length=hop)
    D_direct = scipy.fft.rfft(frames * win[:, np.newaxis], axis=0)
    if dtype == np.float64:
        assert np.allclose(D_direct, D)                      # the reference's assertion, default tolerances
    else:
        assert np.allclose(D_direct, D, rtol=1e-4, atol=1e-5 * np.abs(D_direct).max())


# ------------------------------------------------------------------ test_core.py:308-314
def test_stft_winsizes(lb):
    x = np.zeros(1000000)
    for power in range(12, 17):
        N = 2 ** power
        D = lb.stft(x, n_fft=N, hop_length=N // 2, win_length=N)
        assert D.shape == (N // 2 + 1, 1 + len(x) // (N // 2)) and D.dtype == np.complex128
        assert not D.any()


def test_large_frames_float32_take_the_fp64_kernels(lb):
    """n_fft the float32 kernels are not built for (2^14, 3001) still work and return librosa's dtype."""
    rng = np.random.default_rng(3)
    y = rng.standard_normal(50000).astype(np.float32)
    for n_fft in (16384, 3001):
        D = lb.stft(y, n_fft=n_fft)
        assert D.dtype == np.complex64
        win = scipy.signal.get_window("hann", n_fft, fftbins=True)
        frames = lb.util.frame(np.pad(y.astype(np.float64), n_fft // 2), frame_length=n_fft, hop_length=n_fft // 4)
        ref = scipy.fft.rfft(frames * win[:, None], axis=0)
        assert np.allclose(D, ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())
        yr = lb.istft(D, n_fft=n_fft, length=len(y))
        assert yr.dtype == np.float32 and np.allclose(yr, y, atol=1e-4)


# ------------------------------------------------------------------ test_core.py:317-390
GRID = [(1023, 128), (1023, 129), (1023, 256), (2048, 512), (2048, 2048)]


@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("n_fft, hop_length", GRID)
@pytest.mark.parametrize("N", [1024, 2048, 8192])
def test_stft_preallocate(lb, center, n_fft, hop_length, N):
    y = np.random.default_rng(N).standard_normal(size=(2, max(N, n_fft)))
    D1 = lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length)
    out = np.empty_like(D1)
    D2 = lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length, out=out)
    assert D2 is out
    assert np.allclose(D1, D2)


@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("n_fft, hop_length", GRID)
def test_stft_preallocate_oversize_and_undersize(lb, center, n_fft, hop_length):
    y = np.random.default_rng(7).standard_normal(size=(2, max(2048, n_fft)))
    D1 = lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length)
    shape = list(D1.shape)
    shape[-1] *= 2
    out = np.empty_like(D1, shape=shape)
    D2 = lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length, out=out)
    assert np.allclose(D1, D2) and np.allclose(D1, out[..., : D2.shape[-1]])
    if D1.shape[-1] > 1:
        shape[-1] = D1.shape[-1] // 2
        with pytest.raises(lb.ParameterError):
            lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length, out=np.empty_like(D1, shape=shape))


@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("n_fft, hop_length", GRID)
@pytest.mark.parametrize("N", [1024, 8192])
def test_istft_preallocate(lb, center, n_fft, hop_length, N):
    y = np.random.default_rng(N + 1).standard_normal(size=(2, max(N, n_fft)))
    D = lb.stft(y, center=center, n_fft=n_fft, hop_length=hop_length)
    y1 = lb.istft(D, center=center, n_fft=n_fft, hop_length=hop_length)
    y2 = np.empty_like(y1)
    y3 = lb.istft(D, center=center, n_fft=n_fft, hop_length=hop_length, out=y2)
    assert y3 is y2 and y1.dtype == np.float64
    assert np.allclose(y1, y2)


# ------------------------------------------------------------------ test_core.py:813-828
@pytest.mark.parametrize("sr", [22050, 44100])
@pytest.mark.parametrize("n_fft", [1024, 1025, 2048, 4096])
@pytest.mark.parametrize("window", ["hann", "blackmanharris"])
@pytest.mark.parametrize("hop_length", [128, 256, 512])
def test_istft_reconstruction(lb, sr, n_fft, hop_length, window):
    x = chirp(sr)
    S = lb.stft(x, n_fft=n_fft, hop_length=hop_length, window=window)
    assert S.dtype == np.complex128
    xr = lb.istft(S, hop_length=hop_length, window=window, n_fft=n_fft, length=len(x))
    assert xr.dtype == np.float64
    assert np.all(np.isfinite(xr))
    assert np.allclose(x, xr, atol=1e-6)


@pytest.mark.parametrize("n_fft", [1024, 2048, 4096])
@pytest.mark.parametrize("hop_length", [128, 256, 512])
def test_istft_reconstruction_float32(lb, n_fft, hop_length):
    """The same grid on the float32 hot path, at the tolerance float32 allows (SNR >= 60 dB is BASELINE's gate)."""
    x = chirp(22050).astype(np.float32)
    S = lb.stft(x, n_fft=n_fft, hop_length=hop_length)
    xr = lb.istft(S, hop_length=hop_length, n_fft=n_fft, length=len(x))
    assert xr.dtype == np.float32 and np.all(np.isfinite(xr))
    assert np.allclose(x, xr, atol=2e-6)
    assert 10 * np.log10(np.sum(x.astype(np.float64) ** 2) / np.sum((x - xr).astype(np.float64) ** 2)) > 100.0


# ------------------------------------------------------------------ test_multichannel.py:96-112, 266-285, 653-714
@pytest.fixture(scope="module", params=[np.float64, np.float32])
def y_multi(request):
    rng = np.random.default_rng(11)
    t = np.arange(30000) / 22050.0
    a = 0.3 * np.sin(2 * np.pi * 330 * t) + 0.1 * rng.standard_normal(t.size)
    b = 0.3 * scipy.signal.chirp(t, 60, t[-1], 5000) + 0.1 * rng.standard_normal(t.size)
    return np.stack([a, b]).astype(request.param), 22050


def test_stft_multi(lb, y_multi):
    y, sr = y_multi
    D = lb.stft(y)
    D0, D1 = lb.stft(y[0]), lb.stft(y[1])
    assert np.allclose(D[0], D0) and np.allclose(D[1], D1)
    assert not np.allclose(D0, D1)


def test_istft_multi(lb, y_multi):
    y, sr = y_multi
    D = lb.stft(y)
    y0m, y1m = lb.istft(D[0]), lb.istft(D[1])
    ys = lb.istft(D)
    assert np.allclose(y0m, ys[0]) and np.allclose(y1m, ys[1])
    assert not np.allclose(ys[0], ys[1])


def test_melspectrogram_multi_time(lb, y_multi):
    y, sr = y_multi
    C0, C1 = lb.feature.melspectrogram(y=y[0]), lb.feature.melspectrogram(y=y[1])
    Call = lb.feature.melspectrogram(y=y)
    assert Call.dtype == y.dtype
    assert np.allclose(C0, Call[0]) and np.allclose(C1, Call[1])
    assert not np.allclose(Call[0], Call[1])


def test_mfcc_multi_time(lb, y_multi):
    y, sr = y_multi
    C0, C1 = lb.feature.mfcc(y=y[0], sr=sr), lb.feature.mfcc(y=y[1], sr=sr)
    Call = lb.feature.mfcc(y=y, sr=sr)
    assert Call.dtype == y.dtype
    assert not np.allclose(Call[0], Call[1])
    assert np.allclose(C0, Call[0]), np.max(np.abs(C0 - Call[0]))
    assert np.allclose(C1, Call[1]), np.max(np.abs(C1 - Call[1]))


def test_melspectrogram_and_mfcc_multi_from_S(lb, y_multi):
    y, sr = y_multi
    S = np.abs(lb.stft(y)) ** 2
    M0, M1, Mall = lb.feature.melspectrogram(S=S[0]), lb.feature.melspectrogram(S=S[1]), lb.feature.melspectrogram(S=S)
    assert np.allclose(M0, Mall[0], rtol=1e-5) and np.allclose(M1, Mall[1], rtol=1e-5)
    A = np.abs(lb.stft(y))
    C0 = lb.feature.mfcc(S=lb.amplitude_to_db(A[0], top_db=None))
    C1 = lb.feature.mfcc(S=lb.amplitude_to_db(A[1], top_db=None))
    Call = lb.feature.mfcc(S=lb.amplitude_to_db(A, top_db=None))
    assert np.allclose(C0, Call[0], atol=1e-3) and np.allclose(C1, Call[1], atol=1e-3)
    assert not np.allclose(Call[0], Call[1])


# ------------------------------------------------------------------ float64 against the oracle (the reference run in double)
@pytest.mark.parametrize("n_fft,hop,pad_mode", [(2048, 512, "constant"), (1024, 256, "reflect"), (400, 160, "edge"),
                                                 (512, 128, "symmetric"), (1000, 250, "linear_ramp")])
def test_float64_matches_oracle(lb, n_fft, hop, pad_mode):
    from oracle import ref_np as O

    y = np.random.default_rng(n_fft).standard_normal((2, 12000)) * 0.1
    D, Do = lb.stft(y, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode), O.stft(y, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode)
    assert D.dtype == Do.dtype == np.complex128
    assert np.allclose(D, Do, rtol=1e-9, atol=1e-11 * np.abs(Do).max())
    yr, yo = lb.istft(Do, hop_length=hop, n_fft=n_fft, length=12000), O.istft(Do, hop_length=hop, n_fft=n_fft, length=12000)
    assert yr.dtype == yo.dtype == np.float64
    assert np.allclose(yr, yo, rtol=1e-9, atol=1e-11)
    M, Mo = lb.feature.melspectrogram(y=y, sr=16000, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode), \
        O.melspectrogram(y=y, sr=16000, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode)
    assert M.dtype == Mo.dtype == np.float64
    assert np.allclose(M, Mo, rtol=1e-9, atol=1e-12 * Mo.max())
    C, Co = lb.feature.mfcc(y=y, sr=16000, n_mfcc=20, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode), \
        O.mfcc(y=y, sr=16000, n_mfcc=20, n_fft=n_fft, hop_length=hop, pad_mode=pad_mode)
    assert C.dtype == Co.dtype == np.float64
    assert np.allclose(C, Co, rtol=1e-8, atol=1e-8)
    P = np.abs(Do) ** 2
    for kw in (dict(), dict(ref=np.max), dict(top_db=None), dict(axes=(-1,)), dict(axes=None, ref=np.max)):
        want = O.power_to_db(P, **kw)
        got = lb.power_to_db(P, **kw)
        assert got.dtype == np.float64 and np.allclose(got, want, rtol=1e-10, atol=1e-9), kw
