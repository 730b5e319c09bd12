"""Where does the end-to-end time of melspectrogram(host batch) go?  Development probe."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench, librosa_b200 as lb
from librosa_b200 import _pipeline as pl, _native as nat
w = bench.WORKLOADS["cfg2"]
host = lb.pinned_empty((w["clips"], w["n"]), np.float32); host[...] = bench.make_batch(w, 0)
kw = w["kw"]
for _ in range(3): out = lb.feature.melspectrogram(y=host, sr=22050, **kw)
ctx = lb.default_context()
# 1) whole call
for k in (1, 2, 4, 8):
    os.environ["B2L_HOST_CHUNKS"] = str(k)
    lb.feature.melspectrogram(y=host, sr=22050, **kw)
    t0 = time.perf_counter()
    for _ in range(5): out = lb.feature.melspectrogram(y=host, sr=22050, **kw)
    print(f"chunks={k}: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
os.environ.pop("B2L_HOST_CHUNKS")
# 2) pieces
L = nat.lib()
d = ctx.empty(host.shape, np.float32)
def timeit(name, fn, reps=5):
    fn(); ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.synchronize(); print(f"{name}: {(time.perf_counter()-t0)/reps*1e3:.2f} ms")
timeit("h2d 903MB", lambda: L.b2l_h2d(ctx.handle, C.c_void_p(d.ptr), host.ctypes.data_as(C.c_void_p), host.nbytes))
o = lb.feature.melspectrogram(y=d, sr=22050, **kw); ho = lb.pinned_empty(o.shape, np.float32)
timeit("d2h 226MB", lambda: L.b2l_d2h(ctx.handle, ho.ctypes.data_as(C.c_void_p), C.c_void_p(o.ptr), ho.nbytes))
timeit("kernel (device resident call)", lambda: lb.feature.melspectrogram(y=d, sr=22050, **kw).free())
t0 = time.perf_counter()
for _ in range(100):
    win, wkey = pl.resolve_window("hann", 2048, 2048); b, bk = pl.mel_basis(22050, 2048, dict(n_mels=128))
print(f"host prep (window+mel basis lookup): {(time.perf_counter()-t0)/100*1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(20): a = lb.pinned_empty((1024, 128, 431), np.float32); del a
print(f"pinned_empty 226MB (pooled): {(time.perf_counter()-t0)/20*1e3:.3f} ms")
# 3) both directions at once on two streams
ctx2 = nat.Context(0); d2 = ctx2.empty(host.shape, np.float32)
def both():
    L.b2l_h2d(ctx.handle, C.c_void_p(d.ptr), host.ctypes.data_as(C.c_void_p), host.nbytes)
    L.b2l_d2h(ctx2.handle, ho.ctypes.data_as(C.c_void_p), C.c_void_p(o.ptr), ho.nbytes)
both(); ctx.synchronize(); ctx2.synchronize(); t0 = time.perf_counter()
for _ in range(5): both()
ctx.synchronize(); ctx2.synchronize(); print(f"h2d 903MB || d2h 226MB: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
