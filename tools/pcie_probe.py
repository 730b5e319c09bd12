"""Measure pinned H2D / D2H bandwidth through the C ABI (context for bench.py's e2e figure)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import librosa_b200 as lb
from librosa_b200 import _native as nat
ctx = lb.default_context()
for mb in (64, 256, 903):
    n = mb * 1000 * 1000 // 4
    h = lb.pinned_empty((n,), np.float32); h[:] = 1.0
    d = ctx.empty((n,), np.float32)
    L = nat.lib()
    for name, fn in (("H2D", lambda: L.b2l_h2d(ctx.handle, C.c_void_p(d.ptr), h.ctypes.data_as(C.c_void_p), h.nbytes)),
                     ("D2H", lambda: L.b2l_d2h(ctx.handle, h.ctypes.data_as(C.c_void_p), C.c_void_p(d.ptr), h.nbytes))):
        fn(); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): fn()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb} MB pinned: {dt*1e3:.2f} ms  {h.nbytes/dt/1e9:.1f} GB/s")
