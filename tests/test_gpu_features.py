"""GPU (-m gpu): the frame-wise consumers of the spectrogram (spectral centroid / bandwidth / rolloff /
flatness, rms, zero-crossing rate) through the public drop-in API, against the oracle on identical inputs
and against fixtures produced by the unmodified reference (tests/golden/features_v1.npz).

Stated tolerances (float32 pipeline vs the reference's float64 FFT):
    spectral_centroid / bandwidth   rtol 1e-4, atol 1e-6 * max|ref|  (Hz)
    spectral_rolloff                equal to the reference bin frequency, except in frames where the
                                    reference's own float32 running sum passes within 2e-5 * total of the
                                    threshold at the deciding bin (there one bin either way is rounding)
    spectral_flatness               rtol 1e-4, atol 1e-7
    rms                             rtol 1e-4, atol 1e-7 * max|ref|
    zero_crossing_rate              exact
    onset_strength(_multi)          rtol 1e-4, atol 1e-3 (dB-domain, as for mfcc)
    spectral_contrast               rtol 1e-4, atol 1e-3 dB (1e-6 * max|ref| with linear=True)
    chroma_stft                     rtol 1e-4, atol 2e-6 (chroma values lie in [0, 1])
    estimate_tuning                 exact (a histogram bin centre)
    decompose.hpss                  rtol 1e-4, atol 1e-6 * max|ref| (identical input on both sides)
    effects.hpss / harmonic / percussive   rtol 1e-4, atol 2e-5 * max|ref| (stft -> masks -> istft)
    reassigned_spectrogram          mags rtol 1e-4, atol 1e-5 * max; freqs / times rtol 1e-4, atol 1e-4 * sr/2 /
                                    1e-4 * n_fft/sr on cells with mag >= 1e-3 * max (see _reassign_close)
    phase_vocoder                   rtol 1e-4, atol 1e-5 * max|ref| (identical input STFT)
    effects.time_stretch            rtol 1e-4, atol 2e-3 * max|ref| (2e-4 * max over the first 1000 samples): the
                                    reference's float32 running phase sum limits reproducibility, see _check
    pcen                            rtol 1e-4, atol 1e-6 * max|ref|
    amplitude_to_db                 rtol 1e-5, atol 1e-4 dB (elementwise on identical input)
    db_to_power / db_to_amplitude   rtol 1e-5
"""
import warnings

import numpy as np
import pytest

from feature_cases import FEATURE_CASES, call, case_args, fixture_names, outputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import librosa_b200

    librosa_b200.default_context()
    return librosa_b200


def _close(got, ref, rtol, atol):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert got.dtype == ref.dtype, (got.dtype, ref.dtype)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


def _rolloff_close(O, case, golden, got, ref):
    """Frames may differ only where the reference's running sum is within rounding of the threshold."""
    assert got.shape == ref.shape and got.dtype == ref.dtype
    args, kw = case_args(case, golden)
    roll = kw.get("roll_percent", 0.85)
    if "S" in kw:
        S, n_fft = kw["S"], 2 * (kw["S"].shape[-2] - 1)
    else:
        n_fft = kw.get("n_fft", 2048)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = O.spectrogram(kw["y"], n_fft=n_fft, hop_length=kw.get("hop_length", 512), power=1,
                              pad_mode=kw.get("pad_mode", "constant"))
    freq = O.fft_frequencies(sr=kw.get("sr", 22050), n_fft=n_fft)
    cum = np.cumsum(S, axis=-2)
    total = cum[..., -1:, :]
    thr = roll * total
    bad = ~np.isclose(got, ref, rtol=1e-6, atol=1e-3)
    if not bad.any():
        return
    # index of the bins chosen by each side; every bin between them must sit within rounding of the threshold
    ig = np.abs(freq[:, None] - np.moveaxis(got, -2, -1)[..., None, :, 0].reshape(-1)[None, :]).argmin(axis=0)
    ir = np.abs(freq[:, None] - np.moveaxis(ref, -2, -1)[..., None, :, 0].reshape(-1)[None, :]).argmin(axis=0)
    cum2 = np.moveaxis(cum, -2, 0).reshape(cum.shape[-2], -1)
    thr2, tot2 = thr.reshape(-1), total.reshape(-1)
    for col in np.flatnonzero(bad.reshape(-1)):
        lo, hi = sorted((ig[col], ir[col]))
        assert hi - lo <= 2, (case["name"], col, lo, hi)
        assert np.all(np.abs(cum2[lo:hi, col] - thr2[col]) <= 2e-5 * tot2[col]), (case["name"], col)
    assert bad.mean() <= 0.05, (case["name"], bad.mean())


@pytest.mark.parametrize("case", FEATURE_CASES, ids=[c["name"] for c in FEATURE_CASES])
def test_feature_case_against_oracle_and_reference_fixture(case, lb, oracle, golden):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gots = outputs(call(lb, case, golden))
        wants = outputs(call(oracle, case, golden))
    fixtures = [golden[k] for k in fixture_names(case, len(gots))]
    assert len(gots) == len(wants) == len(fixtures)
    if case["fn"] == "reassigned_spectrogram":
        _reassign_close(case, gots, wants)
        _reassign_close(case, gots, fixtures)
        return
    for got, want, fixture in zip(gots, wants, fixtures):
        _check(case, golden, oracle, got, want, fixture)


def _reassign_close(case, gots, refs):
    """freqs / times / mags of reassigned_spectrogram.  The reassigned coordinates are ratios of STFT values, so
    they are compared where the reference magnitude is at least 1e-3 of the loudest cell (well above float32
    round-off of the FFT); the NaN pattern must agree except within 5 % of the power threshold."""
    kw = case["kw"]
    sr, n_fft = kw.get("sr", 22050), kw.get("n_fft", 2048)
    f_ref, t_ref, m_ref = refs
    f_got, t_got, m_got = gots
    for a, b in zip(gots, refs):
        assert a.shape == b.shape and a.dtype == b.dtype
    scale = float(m_ref.max())
    np.testing.assert_allclose(m_got, m_ref, rtol=1e-4, atol=1e-5 * scale)
    thr = float(kw.get("ref_power", 1e-6)) ** 0.5
    clear = np.abs(m_ref - thr) > 0.05 * thr
    for got, ref, atol in ((f_got, f_ref, 1e-4 * sr / 2), (t_got, t_ref, 1e-4 * n_fft / sr)):
        assert np.array_equal(np.isnan(got)[clear], np.isnan(ref)[clear])
        sel = (m_ref >= 1e-3 * scale) & ~np.isnan(ref) & ~np.isnan(got)
        assert sel.mean() > 0.2
        np.testing.assert_allclose(got[sel], ref[sel], rtol=1e-4, atol=atol)


def _check(case, golden, oracle, got, want, fixture):
    fn = case["fn"]
    for ref in (want, fixture):
        scale = float(np.abs(ref).max()) if ref.size else 0.0
        if fn in ("spectral_centroid", "spectral_bandwidth"):
            _close(got, ref, 1e-4, 1e-6 * scale)
        elif fn == "spectral_rolloff":
            _rolloff_close(oracle, case, golden, got, ref)
        elif fn == "spectral_flatness":
            _close(got, ref, 1e-4, 1e-7)
        elif fn == "rms":
            _close(got, ref, 1e-4, 1e-7 * scale)
        elif fn in ("onset_strength", "onset_strength_multi"):
            _close(got, ref, 1e-4, 1e-3)          # means of dB differences: same absolute term as mfcc
        elif fn == "spectral_contrast":
            # dB of the mean of a sub-band's quietest bins (alpha = 2 %: one or two bins in the low bands), 1e3-1e4
            # below the frame's loudest bin: float32 round-off of the FFT is 1e-7 of THAT bin, i.e. up to 1e-3 of
            # the valley, 4e-3 dB.  Any reordering of the butterfly arithmetic moves the value that much.
            linear = case["kw"].get("linear", False)
            _close(got, ref, 1e-4, 1e-6 * scale if linear else 1e-2)
        elif fn == "chroma_stft":
            _close(got, ref, 1e-4, 2e-6)           # values in [0, 1]; same tuning on both sides (mix T, see cases)
        elif fn == "estimate_tuning":
            assert got.shape == ref.shape == () and float(got) == pytest.approx(float(ref), abs=1e-12)
        elif fn == "pcen":
            _close(got, ref, 1e-4, 1e-6 * scale)
        elif fn == "amplitude_to_db":
            _close(got, ref, 1e-5, 1e-4)          # same input array on both sides: only log10f rounding
        elif fn in ("db_to_power", "db_to_amplitude"):
            _close(got, ref, 1e-5, 1e-37)
        elif fn == "phase_vocoder":
            _close(got, ref, 1e-4, 1e-5 * scale)  # identical input STFT; float32 running phase sum + atan2f / sincosf
        elif case.get("ns") == "decompose":
            _close(got, ref, 1e-4, 1e-6 * scale)  # identical input array: medians are selections, masks smooth
        elif fn == "resample":
            _close(got, ref, 1e-4, 2e-6 * scale)  # ~25 float32 multiply-adds per sample; SciPy rounds each product
        elif fn == "pitch_shift":
            _close(got, ref, 1e-4, 2e-3 * scale)  # time_stretch's bound (below) carried through the FIR
        elif fn == "time_stretch":
            # The reference keeps the unwrapped phase of every bin as a float32 running sum (np.cumsum); after
            # some tens of frames it reaches hundreds of radians (ulp 3e-5 rad), and two STFTs that differ in
            # the last bit (float32 vs float64 FFT) wrap at different frames, so the sums are rounded
            # differently.  Agreement is therefore limited to ~1e-3 of the peak, growing along the clip; the
            # start of the clip, where the sums are still small, must agree ten times more tightly.
            _close(got, ref, 1e-4, 2e-3 * scale)
            head = slice(0, 1000)
            _close(got[..., head], ref[..., head], 1e-4, 2e-4 * scale)
        elif case.get("ns") == "effects":
            _close(got, ref, 1e-4, 2e-5 * scale)  # stft -> masks -> istft; hard to beat istft's own 1e-5 * max
        else:
            assert got.shape == ref.shape and got.dtype == ref.dtype
            np.testing.assert_array_equal(got, ref)


def test_device_resident_inputs_and_S_path_agree(lb):
    """DeviceArray in -> DeviceArray out; the fused y= kernel and the S= kernel on the stored spectrogram see
    the same magnitudes, so their statistics agree to float32 rounding."""
    import signals

    y = signals.make("B", (3, 2, 20000), seed=11)
    yd = lb.to_device(y)
    Sd, _ = lb._spectrogram(y=yd, n_fft=2048, hop_length=512, power=1)
    for fn, kw in [(lb.feature.spectral_centroid, dict(sr=22050)), (lb.feature.spectral_bandwidth, dict(sr=22050)),
                   (lb.feature.spectral_rolloff, dict(sr=22050)), (lb.feature.spectral_flatness, {})]:
        a = fn(y=yd, **kw)
        b = fn(S=Sd, **kw)
        h = fn(y=y, **kw)
        assert isinstance(a, lb.DeviceArray) and isinstance(b, lb.DeviceArray)
        assert a.shape == (3, 2, 1, 40) and h.shape == (3, 2, 1, 40)
        np.testing.assert_allclose(a.get(), b.get(), rtol=2e-6, atol=1e-6 * float(np.abs(h).max()))
        np.testing.assert_allclose(a.get(), h.astype(np.float32), rtol=1e-6)
    r = lb.feature.rms(y=yd)
    z = lb.feature.zero_crossing_rate(yd)
    assert isinstance(r, lb.DeviceArray) and isinstance(z, lb.DeviceArray)
    np.testing.assert_allclose(r.get(), lb.feature.rms(y=y), rtol=1e-6)
    np.testing.assert_allclose(z.get(), lb.feature.zero_crossing_rate(y), rtol=1e-6)


def test_device_resident_variants_of_the_wider_features(lb):
    """DeviceArray in -> DeviceArray out for the features that compose several kernels; the results equal the
    host-input calls (same kernels, only the staging differs)."""
    import signals

    y = signals.make("T", (2, 30000), seed=9)
    yd = lb.to_device(y)
    D = lb.stft(yd)                                         # complex64, kernels' native layout
    P = lb._spectrogram(y=yd, power=2)[0]
    pairs = [
        (lambda: lb.feature.chroma_stft(y=yd, sr=22050), lambda: lb.feature.chroma_stft(y=y, sr=22050)),
        (lambda: lb.feature.chroma_stft(S=P, sr=22050, tuning=0.1), lambda: lb.feature.chroma_stft(y=y, sr=22050, tuning=0.1)),
        (lambda: lb.feature.spectral_contrast(y=yd, sr=22050), lambda: lb.feature.spectral_contrast(y=y, sr=22050)),
        (lambda: lb.onset.onset_strength_multi(y=yd, sr=22050, channels=[0, 64, 128]),
         lambda: lb.onset.onset_strength_multi(y=y, sr=22050, channels=[0, 64, 128])),
        (lambda: lb.decompose.hpss(D), lambda: lb.decompose.hpss(D.get())),
        (lambda: lb.effects.hpss(yd), lambda: lb.effects.hpss(y)),
        (lambda: lb.phase_vocoder(D, rate=1.25), lambda: lb.phase_vocoder(D.get(), rate=1.25)),
        (lambda: lb.effects.time_stretch(yd, rate=1.25), lambda: lb.effects.time_stretch(y, rate=1.25)),
        (lambda: lb.reassigned_spectrogram(yd, sr=22050, fill_nan=True), lambda: lb.reassigned_spectrogram(y, sr=22050, fill_nan=True)),
    ]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for dev_fn, host_fn in pairs:
            dev, host = dev_fn(), host_fn()
            dev = dev if isinstance(dev, tuple) else (dev,)
            host = host if isinstance(host, tuple) else (host,)
            assert len(dev) == len(host)
            for a, b in zip(dev, host):
                assert isinstance(a, lb.DeviceArray) and a.shape == b.shape
                np.testing.assert_allclose(a.get(), b.astype(a.dtype), rtol=1e-6, atol=1e-6 * float(np.abs(b).max()))
        t_dev = lb.estimate_tuning(S=P, sr=22050)
        t_host = lb.estimate_tuning(S=P.get(), sr=22050)
        assert t_dev == t_host


def test_randomized_configurations_against_the_oracle(lb, oracle):
    """Seeded sweep over transform sizes, hops, centring, padding modes, window specs and leading shapes for the
    frame-wise features (the fixtures pin a few dozen hand-picked cases; this covers the cross product)."""
    import signals

    rng = np.random.default_rng(20260922)
    pads = ["constant", "reflect", "edge", "symmetric"]
    for trial in range(24):
        n_fft = int(rng.choice([256, 512, 1024, 2048, 4096, 400, 1000]))
        hop = int(rng.choice([n_fft // 4, n_fft // 2, n_fft // 8 + 3]))
        center = bool(rng.integers(0, 2))
        shape = [(7 * n_fft + int(rng.integers(0, 999)),), (2, 6 * n_fft + 5), (2, 2, 5 * n_fft)][int(rng.integers(0, 3))]
        mix = "ABT"[int(rng.integers(0, 3))]
        sr = int(rng.choice([16000, 22050, 44100]))
        y = signals.make(mix, shape, seed=100 + trial, sr=sr)
        kw = dict(n_fft=n_fft, hop_length=hop, center=center, pad_mode=pads[int(rng.integers(0, 4))],
                  window=["hann", "hamming", ("tukey", 0.3)][int(rng.integers(0, 3))])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            checks = [
                (lb.feature.spectral_centroid(y=y, sr=sr, **kw), oracle.spectral_centroid(y=y, sr=sr, **kw), 1e-4, 1e-6),
                (lb.feature.spectral_flatness(y=y, **kw), oracle.spectral_flatness(y=y, **kw), 1e-4, 1e-7),
                (lb.feature.rms(y=y, frame_length=n_fft, hop_length=hop, center=center),
                 oracle.rms(y=y, frame_length=n_fft, hop_length=hop, center=center), 1e-4, 1e-7),
                (lb.feature.zero_crossing_rate(y, frame_length=n_fft, hop_length=hop, center=center),
                 oracle.zero_crossing_rate(y, frame_length=n_fft, hop_length=hop, center=center), 0.0, 0.0),
            ]
        for got, want, rtol, atol_rel in checks:
            assert got.shape == want.shape and got.dtype == want.dtype, (trial, kw)
            scale = float(np.abs(want).max()) or 1.0
            atol = 1e-7 if atol_rel == 1e-7 and got.dtype == np.float32 and scale < 1.0 else atol_rel * scale   # flatness: absolute
            np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=f"trial {trial} {kw} {shape}")
        for pp in (2.0, 2.5):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = lb.feature.spectral_bandwidth(y=y, sr=sr, p=pp, **kw)
                want = oracle.spectral_bandwidth(y=y, sr=sr, p=pp, **kw)
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6 * float(np.abs(want).max() or 1.0),
                                       err_msg=f"trial {trial} bandwidth p={pp} {kw}")


def test_known_answers_of_the_reference_suite(lb):
    """Same facts (tests/known_answers.py) through the CUDA path; nothing is skipped."""
    import known_answers

    assert known_answers.run(lb.feature, unsupported=(lb.UnsupportedOnGPU,)) == 36


def test_algebraic_properties_of_the_wider_functions(lb):
    """Size-independent identities (no oracle needed): dB conversions invert each other, PCEN with gain 0 /
    bias 0 / power 1 is the identity, harmonic + percussive = input for unit margins, a phase vocoder at rate 1
    returns its input, a time stretch at rate 1 returns the signal, a pure tone reassigns to its frequency."""
    import signals

    rng = np.random.default_rng(5)
    x = rng.uniform(1e-3, 1e3, size=(3, 64, 50)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(lb.db_to_power(lb.power_to_db(x, top_db=None)), x, rtol=2e-5)
        np.testing.assert_allclose(lb.db_to_amplitude(lb.amplitude_to_db(x, top_db=None)), x, rtol=2e-5)
        np.testing.assert_allclose(lb.amplitude_to_db(x), lb.power_to_db(x ** 2, amin=1e-10), atol=1e-4)
        np.testing.assert_allclose(lb.pcen(x, gain=0.0, bias=0.0, power=1.0), x, rtol=1e-5)
        assert not np.any(lb.onset.onset_strength(S=np.full((40, 30), -3.0, dtype=np.float32)))
        y = signals.make("B", (2, 16000), seed=8)
        D = lb.stft(y, n_fft=1024)
        H, P = lb.decompose.hpss(D)
        np.testing.assert_allclose(H + P, D, rtol=1e-4, atol=1e-5 * float(np.abs(D).max()))
        Hm, Pm = lb.decompose.hpss(np.abs(D), mask=True)
        np.testing.assert_allclose(Hm + Pm, 1.0, atol=1e-5)
        np.testing.assert_allclose(lb.phase_vocoder(D, rate=1.0), D, rtol=1e-4, atol=2e-5 * float(np.abs(D).max()))
        ys = lb.effects.time_stretch(y, rate=1.0, n_fft=1024)
        assert ys.shape == y.shape
        np.testing.assert_allclose(ys[..., 512:-512], y[..., 512:-512], atol=2e-4 * float(np.abs(y).max()))
        sr, tone = 22050, 2000.0
        t = np.arange(22050) / sr
        f, tt, m = lb.reassigned_spectrogram((0.5 * np.sin(2 * np.pi * tone * t)).astype(np.float32), sr=sr, fill_nan=True)
        k = int(round(tone * 2048 / sr))
        assert np.allclose(f[k - 1:k + 2, 5:-5], tone, atol=1.0)        # the three bins of the main lobe
        frames_t = np.arange(f.shape[-1]) * 512 / sr
        assert np.allclose(tt[k, 5:-5], frames_t[5:-5], atol=2e-3)       # a stationary tone does not move in time


def test_rms_from_rectangular_stft_matches_rms_from_samples(lb):
    """The reference's own docstring property (feature/spectral.py:872-879): with a constant window and no
    centering, rms(S=|stft|) equals rms(y=...) frame by frame (Parseval)."""
    import signals

    y = signals.make("A", (4, 30000), seed=2)
    S = np.abs(lb.stft(y, window=np.ones, center=False))
    a = lb.feature.rms(S=S)
    b = lb.feature.rms(y=y, center=False)
    np.testing.assert_allclose(a, b, rtol=2e-5)


def test_pcen_streaming_state_matches_one_shot(lb, oracle, golden):
    """zi / return_zf: filtering two halves with the carried state equals one call (the reference's streaming
    contract, core/spectrum.py:2514-2527), and both match the oracle."""
    S = (golden["mel_16000_1024_stereo_A"] * np.float32(2 ** 31)).astype(np.float32)
    full, zf_full = lb.pcen(S, sr=16000, hop_length=256, return_zf=True)
    a, zf_a = lb.pcen(S[..., :10], sr=16000, hop_length=256, return_zf=True)
    b, zf_b = lb.pcen(S[..., 10:], sr=16000, hop_length=256, zi=zf_a, return_zf=True)
    assert zf_a.shape == S.shape[:-1] + (1,) and full.dtype == np.float64
    np.testing.assert_allclose(np.concatenate([a, b], axis=-1), full, rtol=2e-6)
    np.testing.assert_allclose(zf_b, zf_full, rtol=2e-6)
    want, want_zf = oracle.pcen(S, sr=16000, hop_length=256, return_zf=True)
    np.testing.assert_allclose(full, want, rtol=1e-4, atol=1e-6 * float(want.max()))
    np.testing.assert_allclose(zf_full, want_zf, rtol=1e-4)
    d = lb.pcen(lb.to_device(S), sr=16000, hop_length=256)
    assert isinstance(d, lb.DeviceArray)
    np.testing.assert_allclose(d.get(), full.astype(np.float32), rtol=1e-6)


def test_onset_device_resident_and_full_size(lb):
    """cfg-2 sized batch through onset_strength on the device: mel -> dB -> flux, only the envelope comes back."""
    rng = np.random.default_rng(1)
    y = (0.1 * rng.standard_normal((256, 220500))).astype(np.float32)
    y[:, 110250:110250 + 2048] *= 20.0                      # one loud burst in the middle of every clip
    env = lb.onset.onset_strength(y=lb.to_device(y), sr=22050)
    assert isinstance(env, lb.DeviceArray) and env.shape == (256, 431)
    e = env.get()
    assert np.all(e >= 0) and np.all(e[:, :3] == 0)          # lag + centring shift are zero filled
    peak = e.argmax(axis=-1)
    assert np.all(np.abs(peak - (110250 // 512 + 1)) <= 4)    # the flux peaks where the burst starts
    h = lb.onset.onset_strength(y=y[:4], sr=22050)
    np.testing.assert_allclose(h, e[:4], rtol=1e-5, atol=1e-5)


def test_error_behaviour_on_device(lb):
    import signals

    y = signals.make("A", (6000,), seed=4)
    S = np.abs(lb.stft(y, n_fft=512))
    neg = S.copy()
    neg[3, 2] = -1.0
    for fn in (lb.feature.spectral_centroid, lb.feature.spectral_bandwidth, lb.feature.spectral_rolloff):
        with pytest.raises(lb.ParameterError, match="non-negative"):
            fn(S=neg, sr=22050)
        with pytest.raises(lb.ParameterError, match="real-valued"):
            fn(S=S.astype(np.complex64), sr=22050)
    with pytest.raises(lb.ParameterError, match="non-negative"):
        lb.feature.spectral_flatness(S=neg)
    # a clean call right after a failing one must not see a stale flag
    assert np.isfinite(lb.feature.spectral_centroid(S=S, sr=22050)).all()
    bad = y.copy()
    bad[1234] = np.nan
    with pytest.raises(lb.ParameterError, match="not finite"):
        lb.feature.zero_crossing_rate(bad)
    with pytest.raises(lb.ParameterError, match="not finite"):
        lb.feature.spectral_centroid(y=bad)
    with pytest.raises(lb.UnsupportedOnGPU):
        lb.feature.spectral_centroid(S=S, freq=np.ones_like(S))
    with pytest.raises(lb.UnsupportedOnGPU):
        lb.feature.spectral_bandwidth(S=S, centroid=np.ones((1, S.shape[-1])))


def test_full_size_cfg2_statistics_properties(lb):
    """BASELINE cfg-2 shapes (1024 x 10 s @ 22.05 kHz, 2048/512): one fused launch; properties that do not
    need a CPU pass over 441 k frames."""
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal((1024, 220500))).astype(np.float32)
    yd = lb.to_device(y)
    c = lb.feature.spectral_centroid(y=yd).get()
    bw = lb.feature.spectral_bandwidth(y=yd).get()
    ro = lb.feature.spectral_rolloff(y=yd).get()
    fl = lb.feature.spectral_flatness(y=yd).get()
    assert c.shape == (1024, 1, 431)
    # white noise: centroid near sr/4, bandwidth near sr/(4*sqrt(3)), roll-off at 85 % of the band, flat spectrum
    assert abs(float(c.mean()) - 22050 / 4) < 30 and c.min() > 4000 and c.max() < 7000
    assert abs(float(bw.mean()) - 22050 / (4 * np.sqrt(3))) < 40
    assert abs(float(ro.mean()) - 0.85 * 11025) < 60 and ro.max() <= 11025
    assert 0.4 < fl.min() and fl.max() <= 1.0
    # a sample of clips against the oracle-sized path (host input -> same kernels through the chunked pipeline)
    idx = [0, 511, 1023]
    np.testing.assert_allclose(lb.feature.spectral_centroid(y=y[idx]).astype(np.float32), c[idx], rtol=1e-6)
    # time-domain framings at full size
    r = lb.feature.rms(y=yd).get()
    z = lb.feature.zero_crossing_rate(yd).get()
    assert r.shape == (1024, 1, 431) and z.shape == (1024, 1, 431)
    assert abs(float(r[:, :, 2:-2].mean()) - 0.1) < 1e-3
    assert abs(float(z[:, :, 2:-2].mean()) - 0.5) < 5e-3
