"""Where do the ranks' copies of the replicated SVD state part?  The 700 x 1800 parity case on N ranks (separate
processes) under a few settings; SCGBM_RANKCHECK prints, per tsvd call, the first quantities whose checksums differ
between ranks.   python scripts/gpu_rank_consistency.py <nranks>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import test_gpu_multi as T
from tests.util import principal_angle

def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    from oracle.scgbm_oracle import gbm_sc as o_gbm_sc
    kw = dict(max_iter=40)
    ref = o_gbm_sc(T._data(), 8, **kw)
    acc = ref["_accepted"].astype(bool)
    variants = [("baseline", {}), ("bcast", {"SCGBM_SVD_BCAST": "1"}), ("no left start", {"SCGBM_USTART": "0"}),
                ("rankcheck", {"SCGBM_RANKCHECK": "1"})]
    for name, env in variants:
        for k in ("SCGBM_SVD_BCAST", "SCGBM_USTART", "SCGBM_RANKCHECK"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = time.time()
        res = T._spawn(T._worker, world, kw)
        r = res[0]
        n = min(len(r[3]), len(ref["_LL"]))
        rel = np.abs(r[3][:n] - ref["_LL"][:n]) / np.abs(ref["_LL"][:n])
        same = np.array_equal(r[4][:n], ref["_accepted"][:n])
        dU = max(np.abs(q[7] - r[7]).max() for q in res)
        V = np.vstack([q[8] for q in res])
        print(f"== {name}: {time.time() - t0:.1f}s n_iter {len(r[3])} mask equal {same} | LL dev accepted {rel[acc[:n]].max():.2e} rejected "
              f"{rel[~acc[:n]].max():.2e} (iter {int(np.argmax(rel)) + 1}) | U spread over ranks {dU:.2e} | angle vs oracle U "
              f"{principal_angle(r[7], ref['U']):.2e} V {principal_angle(V, ref['V']):.2e} | passes {r[12]} unconverged {r[10]}", flush=True)

if __name__ == "__main__":
    main()
