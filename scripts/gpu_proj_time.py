"""Timing of the projection kernel K9 alone (scgbm_project through gbm_sc(subset=...))."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scgbm_b200 import gbm_sc
from oracle import scgbm_oracle as O
I, J, sub = 20000, int(os.environ.get("PROJ_CELLS", 20000)), 5000
d = O.sim_data(I, J, 20, alpha_mean=-1.5, beta_mean=-1.5, seed=3)
Y = d["Y"]
rng = np.random.default_rng(0)
idx = rng.choice(Y.shape[1], size=sub, replace=False) + 1
for rep in range(2):
    t = time.time(); out = gbm_sc(Y, 20, subset=idx, profile=True); t_p = time.time() - t
    pp = out["_prof_proj"]
    print(f"projection {Y.shape} subset {sub}: total {t_p:.2f}s; project call: device {pp['total_ms']:.1f} ms, kernel classes {{'proj': {pp['ms']['proj']:.1f}, 'pack': {pp['ms']['pack']:.1f}}}, "
          f"IRLS iterations mean {out['_glm_iters'].mean():.1f} max {out['_glm_iters'].max()}", flush=True)
