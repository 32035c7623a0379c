"""get.se at C3 scale (20k genes x 100k cells, M=20) from the kept factors (no W), timing only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from scgbm_b200 import _lib as L, gbm_sc, get_se
import bench
lib = L.load()
I, J, M = 20000, int(os.environ.get("SE_CELLS", 100000)), 20
wl = bench.Workload(I, J, M)
Y = wl.host_slice(lib, L, 0, J, 0)
t = time.time(); out = gbm_sc(Y, M, return_W=False, keep_factors=True, max_iter=10, profile=True); print("fit (10 iterations)", round(time.time() - t, 2), "s")
t = time.time(); se = get_se(out, EPS=0.001, profile=True); dt = time.time() - t
print(f"get.se {I}x{J} M={M} from factors: {dt:.2f}s, kernel classes {se['_prof_se']['ms']['se']:.1f} ms, se_scores median {np.median(se['se_scores']):.4f}")
