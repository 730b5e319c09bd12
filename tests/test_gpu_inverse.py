"""GPU: librosa.feature.inverse on the CUDA path — the reference's own assertions (tests/test_features.py:897-1010)
plus parity with the oracle where the result is unique (mfcc_to_mel) and a residual no worse than the
reference's L-BFGS-B solution where it is not (mel_to_stft: an under-determined NNLS)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import librosa_b200

    librosa_b200.default_context()
    return librosa_b200


def tone(freq=440.0, sr=22050, duration=1.0):
    return np.cos(2 * np.pi * freq * np.arange(int(duration * sr)) / sr - np.pi * 0.5)   # librosa.tone (phi = -pi/2)


@pytest.mark.parametrize("power", [1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_fft", [1024, 2048])
def test_mel_to_stft(lb, oracle, power, dtype, n_fft):
    rng = np.random.default_rng(n_fft + power)
    mel_basis = lb.filters.mel(sr=22050, n_fft=n_fft, n_mels=128, dtype=dtype)
    stft_orig = rng.standard_normal(size=(n_fft // 2 + 1, 4)) ** power
    mels = mel_basis.dot(stft_orig.astype(dtype))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stft = lb.feature.inverse.mel_to_stft(mels, power=power, n_fft=n_fft)
    assert stft.dtype == dtype                       # the reference's four assertions
    assert np.all(stft >= 0)
    assert stft.shape[0] == 1 + n_fft // 2
    rmse = np.sqrt(np.mean((mel_basis.dot(stft ** power) - mels) ** 2))
    assert rmse <= 5e-2
    ref = oracle.mel_to_stft(mels, n_fft=n_fft, power=power)        # the reference's minimiser (L-BFGS-B)
    rmse_ref = np.sqrt(np.mean((mel_basis.dot(ref ** power) - mels) ** 2))
    assert rmse <= rmse_ref * (1 + 1e-3) + 1e-6, (rmse, rmse_ref)


def test_mel_to_stft_batch_and_device(lb):
    rng = np.random.default_rng(4)
    basis = lb.filters.mel(sr=16000, n_fft=512, n_mels=40)
    S = np.abs(rng.standard_normal((3, 2, 257, 9))).astype(np.float32) ** 2
    M = np.einsum("mf,...ft->...mt", basis, S)
    X = lb.feature.inverse.mel_to_stft(M, sr=16000, n_fft=512)
    assert X.shape == S.shape and np.all(X >= 0)
    for i in range(3):                                # batch == per item
        np.testing.assert_array_equal(X[i], lb.feature.inverse.mel_to_stft(M[i], sr=16000, n_fft=512))
    Xd = lb.feature.inverse.mel_to_stft(lb.to_device(M), sr=16000, n_fft=512)
    np.testing.assert_array_equal(Xd.get(), X)
    assert np.sqrt(np.mean((np.einsum("mf,...ft->...mt", basis, X ** 2) - M) ** 2)) <= 1e-3 * M.max()


def test_mel_to_audio(lb):
    y = tone().astype(np.float32)
    M = lb.feature.melspectrogram(y=y, sr=22050)
    y_inv = lb.feature.inverse.mel_to_audio(M, sr=22050, length=len(y))
    assert len(y) == len(y_inv) and y_inv.dtype == np.float32
    assert lb.util.valid_audio(y_inv)
    # the inversion keeps the tone: its mel spectrogram peaks in the same band
    M2 = lb.feature.melspectrogram(y=y_inv, sr=22050)
    assert np.argmax(M2.mean(axis=-1)) == np.argmax(M.mean(axis=-1))


@pytest.mark.parametrize("n_mfcc", [13, 20])
@pytest.mark.parametrize("n_mels", [64, 128])
@pytest.mark.parametrize("dct_type", [2, 3])
@pytest.mark.parametrize("lifter", [-1, 0, 1, 2, 3])
def test_mfcc_to_mel(lb, oracle, n_mfcc, n_mels, dct_type, lifter):
    y = tone().astype(np.float32)
    mfcc = lb.feature.mfcc(y=y, sr=22050, n_mels=n_mels, n_mfcc=n_mfcc, dct_type=dct_type)
    if lifter < 0:
        with pytest.raises(lb.ParameterError):
            lb.feature.inverse.mfcc_to_mel(mfcc * 10 ** 3, n_mels=n_mels, dct_type=dct_type, lifter=lifter)
    elif lifter == 0:
        melspec = lb.feature.melspectrogram(y=y, sr=22050, n_mels=n_mels)
        mel_recover = lb.feature.inverse.mfcc_to_mel(mfcc, n_mels=n_mels, dct_type=dct_type)
        assert melspec.shape == mel_recover.shape and np.all(mel_recover >= 0)
        want = oracle.mfcc_to_mel(mfcc, n_mels=n_mels, dct_type=dct_type)
        assert np.allclose(mel_recover, want, rtol=2e-4, atol=1e-6 * want.max())
    elif lifter == 2:
        with pytest.warns((UserWarning, RuntimeWarning)):
            lb.feature.inverse.mfcc_to_mel(mfcc * 10 ** 3, n_mels=n_mels, dct_type=dct_type, lifter=lifter)
    else:
        ones = np.ones(mfcc.shape, dtype=mfcc.dtype)
        idx = np.arange(1, 1 + mfcc.shape[0], dtype=mfcc.dtype)
        lifter_sine = 1 + lifter * 0.5 * np.sin(np.pi * idx / lifter)[:, np.newaxis]
        mel_recover = lb.feature.inverse.mfcc_to_mel(ones * lifter_sine, n_mels=n_mels, dct_type=dct_type, lifter=lifter)
        mel_expected = lb.feature.inverse.mfcc_to_mel(ones, n_mels=n_mels, dct_type=dct_type, lifter=0)
        np.testing.assert_almost_equal(mel_recover, mel_expected, 3)


@pytest.mark.parametrize("dct_type", [2, 3])
@pytest.mark.parametrize("lifter", [0, 3])
def test_mfcc_to_audio(lb, dct_type, lifter):
    y = tone().astype(np.float32)
    mfcc = lb.feature.mfcc(y=y, sr=22050, n_mels=64, n_mfcc=13, dct_type=dct_type)
    y_inv = lb.feature.inverse.mfcc_to_audio(mfcc, n_mels=64, dct_type=dct_type, lifter=lifter, length=len(y))
    assert len(y) == len(y_inv)
    assert lb.util.valid_audio(y_inv)
