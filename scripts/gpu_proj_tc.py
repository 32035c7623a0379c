"""K9 timing, exact fp64 kernel against the tensor-core path, on the benchmark generator's counts (20k genes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scgbm_b200 import _lib as L, gbm_sc
J = int(os.environ.get("PROJ_CELLS", 60000)); sub = int(os.environ.get("PROJ_SUB", 8000))
cfg = dict(bench.CONFIGS["c4"], J=J)
Y = bench.host_counts(L.load(), L, cfg, 0, J, 0)[0]
idx = np.arange(0, sub) * (J // sub) + 1
res = {}
for method in ("exact", "tcgen05", "tcgen05"):
    t = time.time(); out = gbm_sc(Y, 20, subset=idx, profile=True, glm_method=method); t_p = time.time() - t
    pp = out["_prof_proj"]
    ker = {k: round(v, 1) for k, v in pp["ms"].items() if v > 0}
    print(f"{method}: {Y.shape} subset {sub} genes kept {len(out['_ixs'])}: total {t_p:.2f}s; project call device {pp['total_ms']:.1f} ms, classes {ker}, "
          f"iters mean {out['_glm_iters'].mean():.2f} max {out['_glm_iters'].max()}, fails {int(out['_fail'].sum())}, tc {out['_glm_tc']}", flush=True)
    res[method] = out
a, b = res["exact"], res["tcgen05"]
print("scores max abs diff", np.abs(a["scores"] - b["scores"]).max(), "of scale", np.abs(a["scores"]).max(),
      "| se rel", np.nanmax(np.abs(b["se_scores"] / a["se_scores"] - 1)), "| beta", np.abs(a["beta"] - b["beta"]).max())
