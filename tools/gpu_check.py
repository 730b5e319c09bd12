"""Quick GPU-side parity + timing probe (development aid; the real suite is tests/ -m gpu)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import librosa_b200 as lb
from oracle import ref_np as O

warnings.simplefilter("ignore")
rng = np.random.default_rng(0)

def report(name, got, ref, rtol=1e-4, atol_rel=1e-5):
    got = np.asarray(got); ref = np.asarray(ref)
    ok_shape = got.shape == ref.shape and got.dtype == ref.dtype
    if got.shape != ref.shape:
        print(f"FAIL {name}: shape {got.shape} vs {ref.shape}"); return False
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    ok = np.allclose(got, ref, rtol=rtol, atol=atol_rel * scale) and ok_shape
    print(f"{'ok  ' if ok else 'FAIL'} {name:58s} shape={got.shape} dtype={got.dtype} maxerr={err:.3e} rel_to_max={err/(scale+1e-30):.2e}")
    return ok

allok = True
SECTIONS = sys.argv[1:] or ["parity", "time"]
if "parity" not in SECTIONS:
    cases = []
else:
  cases = [(22050, 2048, 512, True, 'constant'), (5000, 1024, 256, True, 'reflect'), (4000, 512, None, False, 'constant'),
         (1000, 2048, 512, True, 'constant'), (3000, 256, 64, True, 'edge'), (7000, 1024, 300, True, 'symmetric'),
         (6000, 256, 64, True, 'linear_ramp'), (9000, 4096, 1024, True, 'constant'), (3000, 64, 16, True, 'reflect'),
         (3000, 32, 8, True, 'constant'), (2000, 16, 4, True, 'constant'), (2000, 8, 2, False, 'constant'),
         (2000, 128, 37, True, 'constant'), (50000, 2048, 512, False, 'constant'), (33333, 2048, 511, True, 'reflect')]
for n, n_fft, hop, center, pm in cases:
    y = (0.1 * rng.standard_normal(n)).astype(np.float32)
    kw = dict(n_fft=n_fft, hop_length=hop, center=center, pad_mode=pm)
    try:
        D = lb.stft(y, **kw)
        Do = O.stft(y, **kw)
        allok &= report(f"stft n={n} n_fft={n_fft} hop={hop} c={center} {pm}", D, Do)
        yi = lb.istft(Do, hop_length=hop, n_fft=n_fft, center=center)
        yo = O.istft(Do, hop_length=hop, n_fft=n_fft, center=center)
        allok &= report("   istft", yi, yo)
        yi = lb.istft(Do, hop_length=hop, n_fft=n_fft, center=center, length=n)
        yo = O.istft(Do, hop_length=hop, n_fft=n_fft, center=center, length=n)
        allok &= report("   istft length=n", yi, yo)
    except Exception as e:
        import traceback; traceback.print_exc(); allok = False

y2 = (0.1 * rng.standard_normal((2, 3, 30000))).astype(np.float32)
if "parity" not in SECTIONS:
    print("parity skipped")
else:
  exec(compile('''
allok &= report("stft nd", lb.stft(y2, n_fft=1024), O.stft(y2, n_fft=1024))
allok &= report("istft nd", lb.istft(O.stft(y2, n_fft=1024)), O.istft(O.stft(y2, n_fft=1024)))
for sr, n_fft, hop in [(22050, 2048, 512), (16000, 1024, 256), (44100, 4096, 1024), (22050, 512, 128)]:
    allok &= report(f"mel sr={sr} n_fft={n_fft}", lb.feature.melspectrogram(y=y2, sr=sr, n_fft=n_fft, hop_length=hop),
                    O.melspectrogram(y=y2, sr=sr, n_fft=n_fft, hop_length=hop), atol_rel=1e-7)
    allok &= report(f"mfcc sr={sr} n_fft={n_fft}", lb.feature.mfcc(y=y2, sr=sr, n_mfcc=40, n_fft=n_fft, hop_length=hop),
                    O.mfcc(y=y2, sr=sr, n_mfcc=40, n_fft=n_fft, hop_length=hop), rtol=1e-4, atol_rel=1e-5)
allok &= report("mel power=1 htk", lb.feature.melspectrogram(y=y2, sr=22050, power=1.0, htk=True, n_mels=40),
                O.melspectrogram(y=y2, sr=22050, power=1.0, htk=True, n_mels=40), atol_rel=1e-7)
allok &= report("mfcc lifter dct3", lb.feature.mfcc(y=y2, sr=22050, n_mfcc=13, lifter=22, dct_type=3),
                O.mfcc(y=y2, sr=22050, n_mfcc=13, lifter=22, dct_type=3))
Sp = (np.abs(rng.standard_normal((2, 128, 50))) ** 2).astype(np.float32)
allok &= report("power_to_db", lb.power_to_db(Sp), O.power_to_db(Sp), atol_rel=1e-6)
allok &= report("power_to_db ref=max", lb.power_to_db(Sp, ref=np.max), O.power_to_db(Sp, ref=np.max), atol_rel=1e-6)
Sf = (np.abs(rng.standard_normal((2, 1025, 40))) ** 2).astype(np.float32)
allok &= report("mel(S=)", lb.feature.melspectrogram(S=Sf, sr=22050), O.melspectrogram(S=Sf, sr=22050), atol_rel=1e-7)
allok &= report("mfcc(S=)", lb.feature.mfcc(S=lb.power_to_db(Sp), n_mfcc=20), O.mfcc(S=O.power_to_db(Sp), n_mfcc=20), atol_rel=1e-6)
S1, _ = lb._spectrogram(y=y2, n_fft=1024, hop_length=256, power=2.0)
allok &= report("_spectrogram", S1, O.spectrogram(y2, n_fft=1024, hop_length=256, power=2.0), atol_rel=1e-7)
print("ALL OK" if allok else "SOME FAILED")
''', "parity2", "exec"))
if "time" not in SECTIONS:
    sys.exit(0)

# ---- timing: cfg2 (1024 clips x 10 s @ 22050) device-resident
ctx = lb.default_context()
for name, B, n, kw, fn in [
    ("mel cfg2", 1024, 220500, dict(sr=22050, n_fft=2048, hop_length=512), lambda d, kw: lb.feature.melspectrogram(y=d, **kw)),
    ("stft 2048/512", 1024, 220500, dict(n_fft=2048, hop_length=512), lambda d, kw: lb.stft(d, **kw)),
    ("mfcc cfg4-ish (512 clips)", 512, 480000, dict(sr=16000, n_mfcc=40, n_fft=1024, hop_length=256), lambda d, kw: lb.feature.mfcc(y=d, **kw)),
    ("stft 4096/1024 stereo44k (256x2)", 512, 441000, dict(n_fft=4096, hop_length=1024), lambda d, kw: lb.stft(d, **kw)),
]:
    Y = (0.1 * np.random.default_rng(1).standard_normal((B, n))).astype(np.float32)
    d = ctx.to_device(Y)
    out = fn(d, kw); ctx.synchronize()
    T = out.shape[-1]
    e0, e1 = ctx.event(), ctx.event()
    reps = 5
    e0.record()
    for _ in range(reps):
        o = fn(d, kw)
        o.free()
    e1.record()
    ms = e0.elapsed_ms(e1) / reps
    frames = B * T
    print(f"TIME {name:36s} {ms:8.3f} ms  frames={frames}  {frames/ms*1e3/1e6:9.2f} Mframes/s")
    out.free(); d.free()
    if name.startswith("stft 2048"):
        D = lb.stft(d2 := ctx.to_device(Y[:256]), **kw)
        yr = lb.istft(D, hop_length=512, length=n); ctx.synchronize()
        e0.record()
        for _ in range(reps):
            yr2 = lb.istft(D, hop_length=512, length=n); yr2.free()
        e1.record()
        ms = e0.elapsed_ms(e1) / reps
        print(f"TIME istft 2048/512 (256 clips)            {ms:8.3f} ms  frames={256*T}  {256*T/ms*1e3/1e6:9.2f} Mframes/s")
        yh = yr.get(); snr = 10*np.log10((Y[:256]**2).sum() / ((Y[:256]-yh)**2).sum())
        print(f"round-trip SNR {snr:.1f} dB")
