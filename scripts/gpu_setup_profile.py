#!/usr/bin/env python
"""Per-kernel-class CUDA-event breakdown of the setup phase (statistics, init, cold-start SVD) of a fit
on device-generated counts.   python scripts/gpu_setup_profile.py [cells] [factors]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scgbm_b200 import _lib as L

J = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 20
I = 20000
lib = L.load()
cfg = dict(bench.CONFIGS["c5p"], I=I, J=J, M=M, d=M)
ldI = (I + 63) // 64 * 64
ydev = C.c_void_p()
L.check(lib.scgbm_dev_alloc(C.byref(ydev), ldI * J * 2, 0))
bench.sim_into_device(lib, L, cfg, ydev, 0, J, ldI, L.U16, 0)
a = L.FitArgs()
a.Y = ydev; a.y_dtype = L.U16; a.y_on_device = 1; a.I = I; a.J = J; a.ldY = ldI
a.M = M; a.max_iter = 5; a.tol = 1e-4; a.min_iter = 10 ** 6; a.nbatch = 1; a.y_storage = L.STORE_U16; a.device = 0
a.profile = 1
h = C.c_void_p()
t0 = time.perf_counter()
L.check(lib.scgbm_fit_begin(C.byref(a), C.byref(h)))
wall = time.perf_counter() - t0
p = L.Profile()
L.check(lib.scgbm_fit_profile(h, C.byref(p)))
d = p.as_dict()
print(f"setup wall {wall:.3f}s, device total {d['total_ms']:.1f} ms, svd passes {d['svd_passes']}")
print({k: round(v, 1) for k, v in d["ms"].items() if v}, {k: v for k, v in d["launches"].items() if v})
L.check(lib.scgbm_fit_end(h, None))
