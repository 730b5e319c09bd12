"""Parity cases for the frame-wise consumers of the spectrogram: spectral centroid /
bandwidth / rolloff / flatness, rms, zero-crossing rate.  Same three uses as tests/cases.py:
tools/make_golden.py -> tests/golden/features_v1.npz (unmodified reference), the CPU oracle test, the GPU
parity test.

Inputs: ``mix``/``shape`` -> tests/signals.py signal; ``src`` -> |golden stft| of the named hot-path case
(the S= form); ``freq`` -> a synthetic 1-D frequency table of that many bins.
"""
from __future__ import annotations

import numpy as np

FEATURE_CASES = [
    # ---- spectral_centroid
    dict(name="centroid_2048_A", fn="spectral_centroid", mix="A", shape=(9000,), kw=dict(sr=22050)),
    dict(name="centroid_1024_stereo_B", fn="spectral_centroid", mix="B", shape=(2, 6000), kw=dict(sr=16000, n_fft=1024, hop_length=256)),
    dict(name="centroid_C_silence", fn="spectral_centroid", mix="C", shape=(9000,), kw=dict(sr=22050)),
    dict(name="centroid_4096_reflect_B", fn="spectral_centroid", mix="B", shape=(14000,), kw=dict(sr=44100, n_fft=4096, hop_length=1024, pad_mode="reflect")),
    dict(name="centroid_400_nonpow2_A", fn="spectral_centroid", mix="A", shape=(2, 8000), kw=dict(sr=16000, n_fft=400, hop_length=160)),
    dict(name="centroid_fromS", fn="spectral_centroid", src="stft_2048_512_A", kw=dict(sr=22050)),
    dict(name="centroid_fromS_stereo", fn="spectral_centroid", src="stft_512_stereo_A", kw=dict(sr=22050)),
    dict(name="centroid_custom_freq", fn="spectral_centroid", mix="B", shape=(6000,), freq=513, kw=dict(sr=22050, n_fft=1024, hop_length=256)),
    dict(name="centroid_64_16_B", fn="spectral_centroid", mix="B", shape=(2000,), kw=dict(sr=8000, n_fft=64, hop_length=16)),
    # ---- spectral_bandwidth
    dict(name="bandwidth_2048_A", fn="spectral_bandwidth", mix="A", shape=(9000,), kw=dict(sr=22050)),
    dict(name="bandwidth_2048_B", fn="spectral_bandwidth", mix="B", shape=(9000,), kw=dict(sr=22050)),
    dict(name="bandwidth_p3_nonorm_B", fn="spectral_bandwidth", mix="B", shape=(6000,), kw=dict(sr=22050, n_fft=1024, hop_length=256, p=3, norm=False)),
    dict(name="bandwidth_C_silence", fn="spectral_bandwidth", mix="C", shape=(9000,), kw=dict(sr=22050)),
    dict(name="bandwidth_fromS", fn="spectral_bandwidth", src="stft_1024_256_reflect_B", kw=dict(sr=22050)),
    dict(name="bandwidth_501_nonpow2_A", fn="spectral_bandwidth", mix="A", shape=(3000,), kw=dict(sr=22050, n_fft=501, hop_length=128)),
    # ---- spectral_rolloff
    dict(name="rolloff_2048_A", fn="spectral_rolloff", mix="A", shape=(9000,), kw=dict(sr=22050)),
    dict(name="rolloff_095_B", fn="spectral_rolloff", mix="B", shape=(9000,), kw=dict(sr=22050, roll_percent=0.95)),
    dict(name="rolloff_010_stereo_A", fn="spectral_rolloff", mix="A", shape=(2, 6000), kw=dict(sr=16000, n_fft=1024, hop_length=256, roll_percent=0.1)),
    dict(name="rolloff_C_silence", fn="spectral_rolloff", mix="C", shape=(9000,), kw=dict(sr=22050)),
    dict(name="rolloff_fromS", fn="spectral_rolloff", src="stft_2048_512_A", kw=dict(sr=22050, roll_percent=0.5)),
    # ---- spectral_flatness
    dict(name="flatness_2048_A", fn="spectral_flatness", mix="A", shape=(9000,), kw=dict()),
    dict(name="flatness_2048_B", fn="spectral_flatness", mix="B", shape=(9000,), kw=dict()),
    dict(name="flatness_power1_amin_B", fn="spectral_flatness", mix="B", shape=(6000,), kw=dict(n_fft=1024, hop_length=256, power=1.0, amin=1e-6)),
    dict(name="flatness_power3_A", fn="spectral_flatness", mix="A", shape=(4000,), kw=dict(n_fft=512, hop_length=128, power=3.0)),
    dict(name="flatness_C_silence", fn="spectral_flatness", mix="C", shape=(9000,), kw=dict()),
    dict(name="flatness_fromS", fn="spectral_flatness", src="stft_4096_1024_A", kw=dict()),
    # ---- rms
    dict(name="rms_y_A", fn="rms", mix="A", shape=(9000,), ykw=True, kw=dict()),
    dict(name="rms_y_reflect_stereo_B", fn="rms", mix="B", shape=(2, 6000), ykw=True, kw=dict(frame_length=1000, hop_length=250, pad_mode="reflect")),
    dict(name="rms_y_nocenter_C", fn="rms", mix="C", shape=(9000,), ykw=True, kw=dict(frame_length=512, hop_length=100, center=False)),
    dict(name="rms_y_edge_odd_A", fn="rms", mix="A", shape=(5000,), ykw=True, kw=dict(frame_length=1001, hop_length=333, pad_mode="edge")),
    dict(name="rms_S_2048", fn="rms", src="stft_2048_512_A", kw=dict(frame_length=2048)),
    dict(name="rms_S_odd_1025", fn="rms", src="stft_1025_nonpow2", kw=dict(frame_length=1025)),
    # ---- zero_crossing_rate
    dict(name="zcr_A", fn="zero_crossing_rate", mix="A", shape=(9000,), pos=True, kw=dict()),
    dict(name="zcr_B_nocenter_stereo", fn="zero_crossing_rate", mix="B", shape=(2, 6000), pos=True, kw=dict(frame_length=1000, hop_length=300, center=False)),
    dict(name="zcr_C_threshold", fn="zero_crossing_rate", mix="C", shape=(9000,), pos=True, kw=dict(threshold=1e-3)),
    dict(name="zcr_sign_form_pad_C", fn="zero_crossing_rate", mix="C", shape=(9000,), pos=True, kw=dict(zero_pos=False, pad=True, frame_length=512, hop_length=128)),
    dict(name="zcr_ref_magnitude_A", fn="zero_crossing_rate", mix="A", shape=(5000,), pos=True, kw=dict(threshold=0.5, ref_magnitude=0.1, frame_length=256, hop_length=64)),
]

FEATURE_CASES += [
    # ---- dB conversions (top-level functions): amplitude_to_db, db_to_power, db_to_amplitude
    dict(name="amp_to_db_stft", fn="amplitude_to_db", ns="top", arg="stft_2048_512_A", arg_op="abs", kw=dict()),
    dict(name="amp_to_db_refmax_stereo", fn="amplitude_to_db", ns="top", arg="stft_512_stereo_A", arg_op="abs", kw=dict(ref=np.max)),
    dict(name="amp_to_db_top40_amin", fn="amplitude_to_db", ns="top", arg="stft_2048_C_burst", arg_op="abs", kw=dict(top_db=40.0, amin=1e-3, ref=2.0)),
    dict(name="amp_to_db_notop", fn="amplitude_to_db", ns="top", arg="stft_64_16_B", arg_op="abs", kw=dict(top_db=None)),
    dict(name="db_to_power_default", fn="db_to_power", ns="top", arg="const/power_to_db_out", kw=dict()),
    dict(name="db_to_power_ref", fn="db_to_power", ns="top", arg="mfcc_16000_1024_A", kw=dict(ref=3.5)),
    dict(name="db_to_amplitude_default", fn="db_to_amplitude", ns="top", arg="const/power_to_db_out", kw=dict()),
    dict(name="db_to_amplitude_ref", fn="db_to_amplitude", ns="top", arg="const/power_to_db_out_top40", kw=dict(ref=0.25)),
]

FEATURE_CASES += [
    # ---- onset strength (librosa.onset): default mel / dB / flux, sub-bands, max filter, detrend, S= input
    dict(name="onset_default_A", fn="onset_strength", ns="onset", mix="A", shape=(20000,), kw=dict(sr=22050)),
    dict(name="onset_default_C_burst", fn="onset_strength", ns="onset", mix="C", shape=(20000,), kw=dict(sr=22050)),
    dict(name="onset_stereo_lag2_max3_B", fn="onset_strength", ns="onset", mix="B", shape=(2, 16000), kw=dict(sr=16000, lag=2, max_size=3, n_fft=1024, hop_length=256)),
    dict(name="onset_detrend_nocenter_C", fn="onset_strength", ns="onset", mix="C", shape=(20000,), kw=dict(sr=22050, detrend=True, center=False)),
    dict(name="onset_max4_nmels64_A", fn="onset_strength", ns="onset", mix="A", shape=(12000,), kw=dict(sr=22050, max_size=4, n_mels=64, fmax=8000.0)),
    dict(name="onset_multi_4bands_C", fn="onset_strength_multi", ns="onset", mix="C", shape=(20000,), kw=dict(sr=22050, channels=[0, 32, 64, 96, 128])),
    dict(name="onset_multi_noagg_B", fn="onset_strength_multi", ns="onset", mix="B", shape=(9000,), kw=dict(sr=22050, aggregate=False, n_mels=40)),
    dict(name="onset_fromS_db", fn="onset_strength", ns="onset", src_db="mel_22050_2048_C", kw=dict(sr=22050)),
]

FEATURE_CASES += [
    # ---- pcen (top level): defaults on a mel power spectrogram scaled as in the reference's docstring
    dict(name="pcen_default", fn="pcen", ns="top", arg="mel_22050_2048_B", arg_op="scale31", kw=dict(sr=22050)),
    dict(name="pcen_stereo_explicit_b", fn="pcen", ns="top", arg="mel_16000_1024_stereo_A", arg_op="scale31", kw=dict(sr=16000, hop_length=256, b=0.1, gain=0.8, bias=10, power=0.25)),
    dict(name="pcen_power0", fn="pcen", ns="top", arg="mel_22050_2048_C", arg_op="scale31", kw=dict(power=0)),
    dict(name="pcen_bias0", fn="pcen", ns="top", arg="mel_22050_2048_A", kw=dict(bias=0, eps=1e-3)),
    dict(name="pcen_max3", fn="pcen", ns="top", arg="mel_22050_2048_B", arg_op="scale31", kw=dict(max_size=3)),
    dict(name="pcen_max2_stereo_maxaxis", fn="pcen", ns="top", arg="mel_16000_1024_stereo_A", arg_op="scale31", kw=dict(max_size=2, max_axis=-2, time_constant=0.1)),
]

FEATURE_CASES += [
    # ---- spectral_contrast
    dict(name="contrast_default_A", fn="spectral_contrast", mix="A", shape=(9000,), kw=dict(sr=22050)),
    dict(name="contrast_default_B", fn="spectral_contrast", mix="B", shape=(9000,), kw=dict(sr=22050)),
    dict(name="contrast_linear_stereo_B", fn="spectral_contrast", mix="B", shape=(2, 6000), kw=dict(sr=16000, n_fft=1024, hop_length=256, linear=True, n_bands=4)),
    dict(name="contrast_q25_fmin100_A", fn="spectral_contrast", mix="A", shape=(6000,), kw=dict(sr=22050, n_fft=1024, hop_length=256, quantile=0.25, fmin=100.0)),
    dict(name="contrast_C_burst", fn="spectral_contrast", mix="C", shape=(9000,), kw=dict(sr=22050)),
    dict(name="contrast_fromS", fn="spectral_contrast", src="stft_4096_1024_A", kw=dict(sr=44100)),
    dict(name="contrast_400_nonpow2_A", fn="spectral_contrast", mix="A", shape=(2, 8000), kw=dict(sr=16000, n_fft=400, hop_length=160, n_bands=5)),
]

FEATURE_CASES += [
    # ---- chroma_stft and estimate_tuning.  Mix T (detuned harmonic notes) gives the residual histogram of the
    # tuning estimate one clear maximum; with noise-like signals its argmax is a coin toss between tied bins.
    dict(name="chroma_default_T", fn="chroma_stft", mix="T", seed=0, shape=(20000,), kw=dict(sr=22050)),
    dict(name="chroma_stereo_one_tuning_T", fn="chroma_stft", mix="T", seed=1, shape=(2, 16000), kw=dict(sr=16000, n_fft=1024, hop_length=256)),
    dict(name="chroma_tuning_given_norm2_24_B", fn="chroma_stft", mix="B", shape=(9000,), kw=dict(sr=22050, tuning=0.13, norm=2, n_chroma=24, octwidth=None)),
    dict(name="chroma_norm_none_A", fn="chroma_stft", mix="A", shape=(6000,), kw=dict(sr=22050, n_fft=1024, hop_length=256, tuning=0.0, norm=None)),
    dict(name="chroma_norm1_C", fn="chroma_stft", mix="C", shape=(9000,), kw=dict(sr=22050, tuning=-0.2, norm=1)),
    dict(name="chroma_fromS_power", fn="chroma_stft", mix="A", shape=(9000,), as_power_S=True, kw=dict(sr=22050, n_fft=2048, hop_length=512, tuning=0.05)),
    dict(name="chroma_400_nonpow2_T", fn="chroma_stft", mix="T", seed=2, shape=(2, 8000), kw=dict(sr=16000, n_fft=400, hop_length=160, tuning=0.1, base_c=False)),
    dict(name="tuning_y_T", fn="estimate_tuning", ns="top", mix="T", seed=5, shape=(20000,), kw=dict(sr=22050)),
    dict(name="tuning_fromS_stereo_T", fn="estimate_tuning", ns="top", mix="T", seed=2, shape=(2, 16000), as_power_S=True, kw=dict(sr=16000, n_fft=1024, hop_length=256)),
    dict(name="tuning_res05_bpo24_T", fn="estimate_tuning", ns="top", mix="T", seed=1, shape=(20000,), kw=dict(sr=22050, resolution=0.05, bins_per_octave=24, fmin=100.0, fmax=3000.0)),
]

FEATURE_CASES += [
    # ---- harmonic / percussive separation (decompose.hpss on spectrograms, effects.hpss on signals)
    dict(name="hpss_stft_default", fn="hpss", ns="decompose", arg="stft_1024_256_reflect_B", kw=dict()),
    dict(name="hpss_mag_stereo_k13_31_margin", fn="hpss", ns="decompose", arg="stft_512_stereo_A", arg_op="abs", kw=dict(kernel_size=(13, 31), margin=(1.0, 3.0))),
    dict(name="hpss_masks_power1", fn="hpss", ns="decompose", arg="stft_2048_C_burst", kw=dict(mask=True, power=1.0)),
    dict(name="hpss_hard_k8", fn="hpss", ns="decompose", arg="stft_64_16_B", arg_op="abs", kw=dict(power=np.inf, kernel_size=8)),
    dict(name="hpss_k63", fn="hpss", ns="decompose", arg="stft_2048_512_A", arg_op="abs", kw=dict(kernel_size=(63, 5), margin=2.0)),
    dict(name="effects_hpss_B", fn="hpss", ns="effects", mix="B", shape=(2, 12000), pos=True, kw=dict(n_fft=1024)),
    dict(name="effects_harmonic_T", fn="harmonic", ns="effects", mix="T", seed=3, shape=(20000,), pos=True, kw=dict(margin=3.0)),
    dict(name="effects_percussive_C", fn="percussive", ns="effects", mix="C", shape=(20000,), pos=True, kw=dict(kernel_size=17)),
]

FEATURE_CASES += [
    # ---- reassigned_spectrogram (top level): three STFTs with the window, its derivative, its time-weighted form
    dict(name="reassign_default_A", fn="reassigned_spectrogram", ns="top", mix="A", shape=(9000,), pos=True, kw=dict(sr=22050)),
    dict(name="reassign_fill_stereo_B", fn="reassigned_spectrogram", ns="top", mix="B", shape=(2, 9000), pos=True, kw=dict(sr=22050, n_fft=1024, hop_length=200, fill_nan=True)),
    dict(name="reassign_freq_only_noclip_A", fn="reassigned_spectrogram", ns="top", mix="A", shape=(6000,), pos=True, kw=dict(sr=16000, n_fft=512, reassign_times=False, clip=False)),
    dict(name="reassign_times_only_nocenter_A", fn="reassigned_spectrogram", ns="top", mix="A", shape=(6000,), pos=True, kw=dict(sr=22050, n_fft=512, center=False, reassign_frequencies=False, ref_power=0.0)),
    dict(name="reassign_1025_hamming_T", fn="reassigned_spectrogram", ns="top", mix="T", seed=4, shape=(12000,), pos=True, kw=dict(sr=22050, n_fft=1025, window="hamming", win_length=800)),
]

FEATURE_CASES += [
    # ---- phase_vocoder (top level, on fixture STFTs) and effects.time_stretch
    dict(name="pv_rate2", fn="phase_vocoder", ns="top", arg="stft_1024_256_reflect_B", kw=dict(rate=2.0)),
    dict(name="pv_rate_half_stereo", fn="phase_vocoder", ns="top", arg="stft_512_stereo_A", kw=dict(rate=0.5)),
    dict(name="pv_rate137_C", fn="phase_vocoder", ns="top", arg="stft_2048_C_burst", kw=dict(rate=1.37)),
    dict(name="pv_t_out", fn="phase_vocoder", ns="top", arg="stft_2048_512_A", kw=dict(t_out=np.array([0.0, 0.5, 3.2, 3.2, 10.9, 17.99]))),
    dict(name="time_stretch_15_B", fn="time_stretch", ns="effects", mix="B", shape=(2, 12000), pos=True, kw=dict(rate=1.5, n_fft=1024)),
    dict(name="time_stretch_07_T", fn="time_stretch", ns="effects", mix="T", seed=6, shape=(20000,), pos=True, kw=dict(rate=0.7)),
]

FEATURE_CASES += [
    # ---- resample (top level, polyphase = scipy.signal.resample_poly) and effects.pitch_shift on top of it
    dict(name="resample_poly_22050_16000_A", fn="resample", ns="top", mix="A", shape=(9000,), pos=True, kw=dict(orig_sr=22050, target_sr=16000, res_type="polyphase")),
    dict(name="resample_poly_44100_16000_stereo_B", fn="resample", ns="top", mix="B", shape=(2, 12000), pos=True, kw=dict(orig_sr=44100, target_sr=16000, res_type="polyphase")),
    dict(name="resample_poly_up_8000_22050_A", fn="resample", ns="top", mix="A", shape=(3000,), pos=True, kw=dict(orig_sr=8000, target_sr=22050, res_type="polyphase")),
    dict(name="resample_poly_nofix_scale_B", fn="resample", ns="top", mix="B", shape=(5001,), pos=True, kw=dict(orig_sr=22050, target_sr=11025, res_type="polyphase", fix=False, scale=True)),
    dict(name="resample_poly_48000_44100_C", fn="resample", ns="top", mix="C", shape=(3, 6000), pos=True, kw=dict(orig_sr=48000, target_sr=44100, res_type="polyphase")),
    dict(name="pitch_shift_up12_B", fn="pitch_shift", ns="effects", mix="B", shape=(2, 12000), pos=True, kw=dict(sr=22050, n_steps=12, res_type="polyphase", n_fft=1024)),
    dict(name="pitch_shift_down12_T", fn="pitch_shift", ns="effects", mix="T", seed=9, shape=(20000,), pos=True, kw=dict(sr=22050, n_steps=-12, res_type="polyphase")),
]

FEATURE_BY_NAME = {c["name"]: c for c in FEATURE_CASES}


def case_args(case, golden):
    """(positional args, keyword args) for ``<namespace>.<fn>``."""
    import signals

    kw = dict(case["kw"])
    if "arg" in case:           # positional array taken from the hot-path fixtures, optionally transformed
        x = golden[case["arg"]]
        x = {"abs": np.abs, None: lambda v: v, "scale31": lambda v: v * np.float32(2 ** 31)}[case.get("arg_op")](x)
        return (x,), kw
    if "src_db" in case:        # S= form of the onset functions: a dB-scaled mel spectrogram
        p = golden[case["src_db"]]
        kw["S"] = (10.0 * np.log10(np.maximum(1e-10, p))).astype(np.float32)
        return (), kw
    if "src" in case:
        kw["S"] = np.abs(golden[case["src"]])
        return (), kw
    y = signals.make(case["mix"], case["shape"], seed=case.get("seed", len(case["name"])), sr=kw.get("sr", 22050))
    if case.get("as_power_S"):  # S= form fed with |stft|**2 of the signal, computed by the flat oracle (float32)
        from oracle import ref_np as O

        spec_kw = {k: kw[k] for k in ("n_fft", "hop_length") if k in kw}
        kw.pop("hop_length", None)
        kw["S"] = O.spectrogram(y, power=2, **spec_kw)
        return (), kw
    if "freq" in case:
        kw["freq"] = (np.linspace(0.0, 1.0, case["freq"]) ** 2 * 9000.0 + 20.0)
    if case.get("pos"):
        return (y,), kw
    kw["y"] = y
    return (), kw


def resolve(root, case):
    """``root`` is the reference / the drop-in package (functions under ``.feature`` or at top level) or the
    flat oracle module."""
    if case.get("ns") == "top":
        return getattr(root, case["fn"])
    if case.get("ns") == "onset":
        return getattr(getattr(root, "onset", root), case["fn"])
    if case.get("ns") in ("decompose", "effects"):     # oracle: flat names decompose_hpss / effects_hpss
        sub = getattr(root, case["ns"], None)
        return getattr(sub, case["fn"]) if sub is not None else getattr(root, f"{case['ns']}_{case['fn']}")
    return getattr(getattr(root, "feature", root), case["fn"])


def call(root, case, golden):
    args, kw = case_args(case, golden)
    return resolve(root, case)(*args, **kw)


def outputs(result):
    """Results as a list of arrays (functions returning tuples are stored / compared element by element)."""
    return [np.asarray(r) for r in result] if isinstance(result, tuple) else [np.asarray(result)]


def fixture_names(case, n):
    return [case["name"]] if n == 1 else [f"{case['name']}#{i}" for i in range(n)]
