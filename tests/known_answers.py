"""Known-answer checks for the frame-wise features, restated from the facts the reference's own test-suite pins
(tests/test_features.py:127-445 of the reference): single-bin spectra, flat spectra, hand-computed flatness,
constant-band rms, periodic sign flips.  `ns` is the namespace under test — the flat oracle module or the
drop-in package's ``feature`` module — so the same facts pin both."""
import warnings

import numpy as np


def fft_frequencies(sr, n_fft):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def run(ns, unsupported=()):
    """Returns the number of checks performed.  ``unsupported``: exception types that mark a configuration the
    implementation refuses (counted as skipped, e.g. time-varying freq on the GPU)."""
    done = 0

    def attempt(fn):
        nonlocal done
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fn()
            done += 1
        except unsupported:
            pass

    # ---- all energy in DFT band 5 of 513: centroid = freq[5], bandwidth = 0 (test_features.py:127-198)
    S1 = np.zeros((513, 3))
    S1[5, :] = 1.0
    f1024 = fft_frequencies(22050, 1024)
    for freq in (None, f1024, 3 * f1024):
        want = (f1024 if freq is None else freq)[5]

        def centroid(freq=freq, want=want):
            assert np.allclose(ns.spectral_centroid(S=S1, freq=freq), want)

        attempt(centroid)
        for norm in (False, True):
            for p in (1, 2):
                def bandwidth(freq=freq, norm=norm, p=p):
                    assert not np.any(ns.spectral_bandwidth(S=S1, freq=freq, norm=norm, p=p))

                attempt(bandwidth)

    def onecol():
        assert ns.spectral_bandwidth(S=S1[:, :1]).shape == (1, 1)

    attempt(onecol)

    # ---- flat spectrum: pre-computed bandwidths, and std of the bin indices (test_features.py:209-226)
    flat = np.ones((1025, 1))

    def flat_bw():
        assert np.isclose(ns.spectral_bandwidth(S=flat, p=1), 2758.93902439, rtol=1e-5)
        assert np.isclose(ns.spectral_bandwidth(S=flat, p=2), 3185.74989294, rtol=1e-5)
        bins = np.arange(1025)
        assert np.isclose(ns.spectral_bandwidth(S=flat, freq=bins), np.std(bins), rtol=1e-5)

    attempt(flat_bw)

    # ---- roll-off of a flat spectrum sits at floor(pct * bins) (test_features.py:232-256)
    ones = np.ones((1025, 3))
    f2048 = fft_frequencies(22050, 2048)
    for pct in (0.25, 0.5, 0.95):
        def rolloff(pct=pct):
            got = ns.spectral_rolloff(S=ones, sr=22050, roll_percent=pct)
            assert np.allclose(got, f2048[int(np.floor(pct * 1025))])

        attempt(rolloff)

    # ---- contrast of a flat spectrum is 0; one spike in the top band (test_features.py:296-320)
    for linear in (False, True):
        def contrast_flat(linear=linear):
            assert np.allclose(ns.spectral_contrast(S=np.ones((1025, 10)), linear=linear), 0.0, atol=1e-5)

        attempt(contrast_flat)

        def contrast_spike(linear=linear):
            S = np.zeros((1025, 1))
            S[512, 0] = 1.0
            got = ns.spectral_contrast(S=S, linear=linear, quantile=0.001)[:, 0]
            want = [0, 0, 0, 0, 0, 1, 0] if linear else [20, 20, 20, 20, 20, 100, 20]
            assert np.allclose(got, want, atol=1e-4)

        attempt(contrast_spike)

    # ---- flatness: hand-computed 3-bin case, all-ones, all-zeros (test_features.py:345-358)
    def flatness():
        S = np.array([[1, 3], [2, 1], [1, 2]])
        assert np.allclose(ns.spectral_flatness(S=S), [[0.7937005259, 0.7075558390]], rtol=1e-5)
        assert np.allclose(ns.spectral_flatness(S=np.ones((1025, 2))), 1.0)
        assert np.allclose(ns.spectral_flatness(S=np.zeros((1025, 2))), 1.0)

    attempt(flatness)

    # ---- rms of an all-ones band spectrum is 1 / sqrt(frame_length) (test_features.py:372-381)
    for n in range(10, 100, 20):
        def rms_const(n=n):
            L = 2 * (n - 1)
            got = ns.rms(S=np.ones((n, 5)), frame_length=L)
            assert np.allclose(got, 1.0 / np.sqrt(L), atol=1e-2)

        attempt(rms_const)

    # ---- a sign flip every `period` samples gives a crossing rate of 2 / period (test_features.py:415-442)
    for period in (32, 8, 2):
        y = np.ones(16384, dtype=np.float32)
        y[::period] = -1
        for frame_length, hop, center in ((513, 128, False), (2049, 256, True)):
            def zcr(y=y, period=period, frame_length=frame_length, hop=hop, center=center):
                z = ns.zero_crossing_rate(y, frame_length=frame_length, hop_length=hop, center=center)
                if center:
                    z = z[:, frame_length // 2: -frame_length // 2]
                assert np.allclose(z, 2.0 / period, rtol=1e-2)

            attempt(zcr)
    return done
