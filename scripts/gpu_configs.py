"""Timing + parity of BASELINE.json's smaller configs on one B200:
C2: simData-style 2k x 10k NB counts, M=10, full fit + get.se (vs the oracle)
C4-lite: projection gbm.sc(subset=...) on a 20k x N matrix, M=20 (timing only)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scgbm_b200 import gbm_sc, get_se
from oracle import scgbm_oracle as O
from tests.util import principal_angle

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which in ("c2", "all"):
    Y = O.sim_data(2000, 10000, 10, seed=1, nb_theta=1.0)["Y"]
    t = time.time(); out = gbm_sc(Y, 10, diagnostics=True, profile=True); t_fit = time.time() - t
    t = time.time(); out = get_se(out, profile=True); t_se = time.time() - t
    print(f"C2 {Y.shape}: fit {t_fit:.2f}s ({out['_n_iter']} iterations), get.se {t_se:.2f}s, se prof {out.get('_prof_se', {}).get('ms')}", flush=True)
    t = time.time(); ref = O.gbm_sc(Y, 10); t_ofit = time.time() - t
    t = time.time(); ref = O.get_se_fast(ref); t_ose = time.time() - t
    n = min(len(ref["_LL"]), len(out["_LL"]))
    print(f"   oracle fit {t_ofit:.1f}s get.se {t_ose:.1f}s | LLrel {np.max(np.abs(ref['_LL'][:n]-out['_LL'][:n])/np.abs(ref['_LL'][:n])):.2e} "
          f"mask_eq {np.array_equal(ref['_accepted'][:n], out['_accepted'][:n])} alpha {np.abs(ref['alpha']-out['alpha']).max():.2e} "
          f"angU {principal_angle(ref['U'], out['U']):.2e} angV {principal_angle(ref['V'], out['V']):.2e} "
          f"se_scores rel {np.max(np.abs(out['se_scores']/ref['se_scores']-1)):.2e} se_loadings rel {np.max(np.abs(out['se_loadings']/ref['se_loadings']-1)):.2e}", flush=True)
if which in ("proj", "all"):
    I, J, sub = 20000, int(os.environ.get("PROJ_CELLS", 20000)), 5000
    rng = np.random.default_rng(0)
    d = O.sim_data(I, J, 20, alpha_mean=-1.5, beta_mean=-1.5, seed=3)
    Y = d["Y"]
    idx = rng.choice(Y.shape[1], size=sub, replace=False) + 1
    t = time.time(); out = gbm_sc(Y, 20, subset=idx, profile=True); t_p = time.time() - t
    print(f"projection {Y.shape} subset {sub}: {t_p:.2f}s, scores {out['scores'].shape}", flush=True)
