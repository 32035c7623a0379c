#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: gbm.sc outer-loop iterations/second (and fit seconds) at 20k genes x 1M
cells, M = 20, on N B200s of one node; plus every other named configuration of BASELINE.json through --config.

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torchrun; default config c5p)
    python bench.py --config c2|c3|c4|c5|c5p|c2se|c3se ...   (the named shapes, one JSON line each)
    python bench.py --impl reference ...                      (CPU arm, rank 0 only, same config keys)

  c5p  20k x 1M, M = 20, full fit              -- the metric's own configuration (default; "C5'" of SURVEY 8d)
  c2   2k x 10k NB (theta = 1), M = 10, full fit + get.se      (BASELINE configs[1])
  c3   20k x 100k ~90 % zeros, M = 20, full fit                (configs[2])
  c4   gbm.sc(subset = 50k) on 20k x 1M, M = 20, cells sharded over the ranks   (configs[3]; value = cells/s)
  c5   20k x 1M, M = 50, full fit                              (configs[4])
  c2se / c3se   get.se alone at the c2 / c3 shape (seconds; lower is better)

A "step" of a fit config is one trip through the loop body of R/fit.R:125-186 over the whole count matrix (alpha
update, fused likelihood pass, accept/revert control, truncated SVD).
`value`   : K steps / device time, Y already resident in HBM (u16), max over ranks.
`e2e`     : a complete default gbm.sc(Y, M) call through the public API with HOST (pinned, int32 = R integer)
            buffers: H2D of Y, setup, initial SVD, all iterations and the D2H of the results inside the timed region.
`roofline`: the fused likelihood pass (K4) -- algorithmic bytes / mean launch time, against MEASURED_PEAKS.json's copy
            bandwidth; `mix_peak` next to it is what a trivial streaming kernel with the pass's own 1 : 2 read : write
            mix reaches on this GPU, measured in the same process (scgbm_dbg_membench).
`cpu_baseline`: the float64 oracle (numpy/scipy restatement of R/fit.R, NOT R) on a bounded slice of the same
            synthetic model, extrapolated linearly in I*J; all host cores (BLAS threads forced, whatever
            OMP_NUM_THREADS the launcher exported).
"""
from __future__ import annotations

import os
import sys

# The CPU arm must use every host core: torch.distributed.run exports OMP_NUM_THREADS=1 to its children, which
# would silently make the reference's BLAS single-threaded.  Must happen before numpy is imported.
if "--impl" in sys.argv and "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import ctypes as C  # noqa: E402
import json  # noqa: E402
import math  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------
# the named configurations (BASELINE.json configs)
# ----------------------------------------------------------------------------------
def zero_fraction(mu_total, sd2=2.0):
    """P(Y = 0) for Y ~ Poisson(exp(z)), z ~ N(mu_total, sd2) (alpha + beta)."""
    x, w = np.polynomial.hermite_e.hermegauss(80)
    z = mu_total + math.sqrt(sd2) * x
    return float(np.sum(w * np.exp(-np.exp(z))) / np.sum(w))


def solve_mean_for_zeros(target=0.90):
    lo, hi = -10.0, 3.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if zero_fraction(mid) > target:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


_SPARSE_MEAN = solve_mean_for_zeros(0.90) / 2            # alpha.mean = beta.mean giving ~90 % zeros

CONFIGS = {
    "c2":   dict(I=2000, J=10000, M=10, d=10, amean=0.0, bmean=0.0, theta=1.0, what="fit+se", ref="configs[1]"),
    "c3":   dict(I=20000, J=100000, M=20, d=20, amean=_SPARSE_MEAN, bmean=_SPARSE_MEAN, theta=0.0, what="fit", ref="configs[2]"),
    "c4":   dict(I=20000, J=1000000, M=20, d=20, amean=_SPARSE_MEAN, bmean=_SPARSE_MEAN, theta=0.0, what="proj", ref="configs[3]", subset=50000),
    "c5":   dict(I=20000, J=1000000, M=50, d=50, amean=_SPARSE_MEAN, bmean=_SPARSE_MEAN, theta=0.0, what="fit", ref="configs[4]"),
    "c5p":  dict(I=20000, J=1000000, M=20, d=20, amean=_SPARSE_MEAN, bmean=_SPARSE_MEAN, theta=0.0, what="fit", ref="metric (20k x 1M, M=20)"),
    "c2se": dict(I=2000, J=10000, M=10, d=10, amean=0.0, bmean=0.0, theta=1.0, what="se", ref="configs[1] (get.se)"),
    "c3se": dict(I=20000, J=100000, M=20, d=20, amean=_SPARSE_MEAN, bmean=_SPARSE_MEAN, theta=0.0, what="se", ref="configs[2] (get.se)"),
}
SEED = 1


def describe(cfg):
    kind = "NB(theta=%g)" % cfg["theta"] if cfg["theta"] > 0 else "Poisson"
    zeros = "" if cfg["amean"] == 0.0 else ", ~90% zeros"
    call = {"fit": "gbm.sc(Y, M=%d)", "fit+se": "gbm.sc(Y, M=%d) + get.se", "se": "get.se(gbm.sc(Y, M=%d))",
            "proj": "gbm.sc(Y, M=%d, subset=50000 column indices)"}[cfg["what"]] % cfg["M"]
    return f"{cfg['I']}x{cfg['J']} {kind} simData-recipe counts (d={cfg['d']}{zeros}), {call}"


def config_dict(name, cfg, world, extra=None):
    """the `config` object: identical keys in the repo arm and in the reference arm"""
    c = {"workload": describe(cfg), "name": name, "baseline_config": cfg["ref"], "genes": cfg["I"], "cells": cfg["J"],
         "M": cfg["M"], "sharding": f"cells over {world} rank(s)"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------
# synthetic input: scgbm_sim_data draws everything on the device (simData recipe, R/simulate.R:24-48)
# ----------------------------------------------------------------------------------
def sim_args(L, cfg, j0, j1, out_dtype, on_device, ld, device):
    a = L.SimDataArgs()
    a.model = 0; a.d = cfg["d"]; a.I = cfg["I"]; a.J = j1 - j0; a.J_total = cfg["J"]; a.j_offset = j0
    a.alpha_mean = cfg["amean"]; a.beta_mean = cfg["bmean"]; a.K = 2.0; a.theta = cfg["theta"]
    a.seed = SEED; a.out_dtype = out_dtype; a.y_on_device = int(on_device); a.ldY = ld; a.device = device
    return a


def sim_into_device(lib, L, cfg, dev_ptr, j0, j1, ld, out_dtype, device):
    a = sim_args(L, cfg, j0, j1, out_dtype, True, ld, device)
    o = L.SimDataOut(); o.Y = dev_ptr
    L.check(lib.scgbm_sim_data(C.byref(a), C.byref(o)))
    return int(o.n_saturated)


def host_counts(lib, L, cfg, j0, j1, device, pinned=False):
    """int32 host matrix of the cells [j0, j1) (what an R user holds), generated on the device chunk by chunk.
    Returns (array, free_fn, is_pinned)."""
    I = cfg["I"]; n = j1 - j0
    pin = None
    if pinned:
        pin = C.c_void_p()
        if lib.scgbm_host_alloc_pinned(C.byref(pin), I * n * 4) != 0:
            log("pinned allocation failed, using pageable host memory"); pin = None
    if pin is not None:
        Y = np.frombuffer((C.c_int32 * (I * n)).from_address(pin.value), dtype=np.int32).reshape((I, n), order="F")
    else:
        Y = np.empty((I, n), dtype=np.int32, order="F")
    chunk = max(1, (512 << 20) // (I * 4))
    dchunk = C.c_void_p()
    L.check(lib.scgbm_dev_alloc(C.byref(dchunk), I * min(chunk, n) * 4, device))
    try:
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            sim_into_device(lib, L, cfg, dchunk, j0 + c0, j0 + c1, I, L.I32, device)
            L.check(lib.scgbm_memcpy_d2h(C.c_void_p(Y.ctypes.data + c0 * I * 4), dchunk, I * (c1 - c0) * 4, device))
    finally:
        lib.scgbm_dev_free(dchunk, device)

    def free():
        if pin is not None:
            lib.scgbm_host_free_pinned(pin)
    return Y, free, pin is not None


def oracle_sample(cfg, n_cells):
    """The CPU arm's input: the same simData recipe and parameters, first n_cells cells, drawn with numpy -- the
    reference arm touches nothing of the repo's GPU code."""
    rng = np.random.default_rng(SEED + 12345)
    I, J, d = cfg["I"], cfg["J"], cfg["d"]
    n = min(n_cells, J)
    U, _ = np.linalg.qr(rng.standard_normal((I, d)))
    V, _ = np.linalg.qr(rng.standard_normal((n, d)))
    V = V * math.sqrt(n / J)                       # rows of a J-cell Haar block have squared norm ~ d / J
    s = math.sqrt(I) + math.sqrt(J)
    D = np.linspace(2 * s, s, d)
    alpha = rng.normal(cfg["amean"], 1.0, I); beta = rng.normal(cfg["bmean"], 1.0, n)
    lam = np.exp(alpha[:, None] + beta[None, :] + (U * D) @ V.T)
    if cfg["theta"] > 0:
        lam = rng.gamma(cfg["theta"], lam / cfg["theta"])
    Y = rng.poisson(lam).astype(np.int32)
    return Y[Y.sum(1) > 0][:, Y.sum(0) > 0]


# ----------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi style clocks line sampled DURING the timed region (pynvml)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log("pynvml unavailable:", e)
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(2.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def torch_broadcast_bytes(nbytes=128, src=0):
    """rendezvous only: hands every rank the NCCL unique id of libscgbm over torch.distributed (gloo)."""
    import torch
    import torch.distributed as dist

    def _bc(payload):
        t = torch.zeros(nbytes, dtype=torch.uint8)
        if payload is not None:
            t = torch.tensor(list(payload), dtype=torch.uint8)
        dist.broadcast(t, src=src)
        return bytes(t.tolist())
    return _bc


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        return ", ".join(sorted({f"{i.get('internal_api')} {i.get('version')} x{i.get('num_threads')}" for i in threadpool_info()}))
    except Exception:
        return "unknown"


def cpu_baseline(cfg, steps, warmup, sample_cells):
    """oracle (numpy/scipy float64 restatement, svds as the irlba stand-in) on a bounded slice of the same synthetic
    model (first `sample_cells` cells, all genes), every host core."""
    from oracle import scgbm_oracle as O
    ncores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncores)       # whatever OMP_NUM_THREADS said when the BLAS was loaded
    except Exception:
        pass
    Ys = oracle_sample(cfg, sample_cells)
    M = cfg["M"]
    t0 = time.perf_counter()
    out = O.gbm_sc(Ys, M, max_iter=warmup + steps, min_iter=10 ** 6, svd="svds", time_by_iter=True,
                   rng=np.random.default_rng(0))
    wall = time.perf_counter() - t0
    t = np.asarray(out["time"])              # cumulative, sampled mid-iteration (R/fit.R:139-142)
    per_iter = np.diff(t)[-max(1, steps - 1):] if len(t) > 2 else np.array([wall / max(1, len(t))])
    sec_iter = float(np.mean(per_iter))
    se_sec = None
    if cfg["what"] in ("fit+se", "se"):
        t1 = time.perf_counter()
        O.get_se_fast(out)
        se_sec = time.perf_counter() - t1
    scale = (Ys.shape[0] * Ys.shape[1]) / (cfg["I"] * cfg["J"])
    blas = blas_info()
    return {
        "value": scale / sec_iter, "unit": "iters/s", "cores": ncores, "kind": "port",
        "sample": f"oracle (float64 numpy/scipy restatement of R/fit.R, svds tol=1e-5 for irlba; NOT R) on "
                  f"{Ys.shape[0]}x{Ys.shape[1]} of the {cfg['I']}x{cfg['J']} model, M={M}, {len(t)} iterations, "
                  f"{sec_iter:.3f} s/iter measured, extrapolated linearly in I*J (x{1 / scale:.0f}); BLAS: {blas}",
        "sec_per_iter_sample": sec_iter, "sample_shape": [int(Ys.shape[0]), int(Ys.shape[1])], "wall_s": wall,
        "extrapolation_factor": 1.0 / scale, "get_se_seconds_sample": se_sec, "blas": blas,
    }


# ----------------------------------------------------------------------------------
class Ranks:
    """one process per GPU: gloo for the rendezvous / barriers / max-over-ranks, NCCL (inside libscgbm) for data"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", 0)); self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.dist = None; self.comm = None

    def init(self, device):
        if self.world == 1:
            return
        # stdout must carry exactly one JSON line.  NCCL prints its "NCCL version ..." banner there at every level
        # from VERSION up (WARN included), and a level can also come from /etc/nccl.conf: an explicit value NCCL
        # does not know ("NONE") switches the banner off; a level the user asked for is sent to stderr instead.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "NONE"
        elif "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        import torch.distributed as dist
        from scgbm_b200.dist import init_comm
        dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
        self.dist = dist
        self.comm = init_comm(self.rank, self.world, device, torch_broadcast_bytes())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_array(self, a):
        if self.dist is None:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)); self.dist.all_reduce(t)
        return t.numpy()

    def close(self):
        if self.comm is not None:
            self.comm.destroy()
        if self.dist is not None:
            self.dist.destroy_process_group()


def mix_peak(lib, device):
    """GB/s of a trivial streaming kernel with the fused pass's read : write mix (1 : 2) and of a plain copy, measured
    here and now: the ceiling this traffic pattern has on this GPU (write-heavy streams run below the copy rate)."""
    try:
        g = (C.c_double * 2)()
        if lib.scgbm_dbg_membench(device, g) != 0:
            return None
        return {"copy_1r1w_gbs": g[0], "mix_1r2w_gbs": g[1]}
    except Exception:
        return None


# ----------------------------------------------------------------------------------
def run_fit(name, cfg, args, R):
    from scgbm_b200 import _lib as L, gbm_sc, get_se
    from scgbm_b200.dist import shard_range
    lib = L.load()
    rank, world, device = R.rank, R.world, R.local_rank
    I, J, M = cfg["I"], cfg["J"], cfg["M"]
    K, W = args.steps, max(args.warmup, 0)
    t_gen = time.perf_counter()
    j0, j1 = shard_range(J, rank, world)
    Jl = j1 - j0
    ldI = (I + 63) // 64 * 64
    ydev = C.c_void_p()
    L.check(lib.scgbm_dev_alloc(C.byref(ydev), ldI * Jl * 2, device))
    nsat = sim_into_device(lib, L, cfg, ydev, j0, j1, ldI, L.U16, device)
    log(f"rank {rank}: generated {I}x{Jl} u16 counts on device in {time.perf_counter() - t_gen:.1f}s, saturated={nsat}")

    # ---------------- value: K steps with Y resident in HBM ----------------------
    a = L.FitArgs()
    a.Y = ydev; a.y_dtype = L.U16; a.y_on_device = 1; a.I = I; a.J = Jl; a.ldY = ldI
    NB = 3                      # extra, fully instrumented steps for the per-kernel breakdown (outside the timed region)
    a.M = M; a.max_iter = W + K + NB + 1; a.tol = 1e-4; a.min_iter = 10 ** 6; a.infer_beta = 0
    a.batch = None; a.nbatch = 1; a.return_W = 0; a.time_by_iter = 0
    # AUTO: the library keeps the resident u16 matrix when some count exceeds 255 and repacks it as u8 otherwise
    # (SCGBM_BENCH_STORAGE=u16 pins the two-byte format: A/B measurements)
    a.y_storage = L.STORE_U16 if os.environ.get("SCGBM_BENCH_STORAGE", "auto") == "u16" else L.STORE_AUTO
    a.device = device; a.comm = R.comm.handle if R.comm else None
    a.profile = 2; a.seed = 0   # timed region: CUDA events around the fused pass only (the roofline kernel)
    h = C.c_void_p()
    t0 = time.perf_counter()
    L.check(lib.scgbm_fit_begin(C.byref(a), C.byref(h)))
    setup_s = time.perf_counter() - t0
    log(f"rank {rank}: setup + initial SVD {setup_s:.2f}s")
    done = C.c_int32(0)
    if W > 0:
        L.check(lib.scgbm_fit_step(h, W, C.byref(done)))
    p0 = L.Profile(); p1 = L.Profile()
    sampler = ClockSampler(device)                          # nvmlInit happens here, outside the timed region
    R.barrier()
    sampler.start()
    L.check(lib.scgbm_fit_profile(h, C.byref(p0)))          # records an event on the launch stream + sync
    L.check(lib.scgbm_fit_step(h, K, C.byref(done)))
    L.check(lib.scgbm_fit_profile(h, C.byref(p1)))
    clocks = sampler.stop()
    R.barrier()
    ms = R.max(p1.total_ms - p0.total_ms)
    d0, d1 = p0.as_dict(), p1.as_dict()
    launches = {k: d1["launches"][k] - d0["launches"][k] for k in d1["launches"]}
    fused_ms = (d1["ms"]["fused"] - d0["ms"]["fused"]) / max(1, launches["fused"])
    fused_bytes = d1["fused_bytes"]
    svd_passes = d1["svd_passes"] - d0["svd_passes"]
    # per-kernel-class breakdown: NB more steps with every class bracketed by events (costs ~0.5 ms per step
    # in event records, which is why the timed region above does without)
    p2 = L.Profile(); p3 = L.Profile()
    L.check(lib.scgbm_fit_set_profile(h, 1))
    L.check(lib.scgbm_fit_profile(h, C.byref(p2)))
    L.check(lib.scgbm_fit_step(h, NB, C.byref(done)))
    L.check(lib.scgbm_fit_profile(h, C.byref(p3)))
    d2, d3 = p2.as_dict(), p3.as_dict()
    kms = {k: (d3["ms"][k] - d2["ms"][k]) * K / NB for k in d3["ms"]}     # scaled to K steps like `launches`
    L.check(lib.scgbm_fit_end(h, None))
    peak, peak_src = measured_peak()
    achieved = fused_bytes / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else None
    traffic, traffic_note = None, "no ncu capture committed for this shape"
    tpath = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ratio = tj["traffic_bytes_per_launch"] / tj["algorithmic_bytes_per_launch"]
            traffic = ratio * fused_bytes
            traffic_note = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of the same kernel: {ratio:.3f}x its algorithmic "
                            f"bytes, captured at {tj.get('shape', '20000x100000')} ({tj.get('source', 'profiles/')})"
                            + ("" if tj.get("cells") == Jl else "; scaled to this shard by the algorithmic bytes"))
        except Exception:
            pass
    mp = mix_peak(lib, device) if rank == 0 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": "fused_tc_kernel<uint16_t> (K4 fused likelihood pass: tcgen05 eta tile, TMA in/out; on shards "
                          "of >= 4 cell tiles per SM the launch also carries the SVD's first product R^T Q0, which "
                          "adds no algorithmic HBM bytes and replaces one 4N-byte pass of prod_tc_kernel)",
                "peak_source": peak_src, "bytes_per_launch": fused_bytes, "ms_per_launch": fused_ms,
                "traffic_note": traffic_note,
                "mix_peak": mp,
                "frac_of_mix_peak": (achieved / mp["mix_1r2w_gbs"]) if (achieved and mp) else None,
                "note": "algorithmic bytes = N*s_Y + 4N (two fp16 residual planes) + 4M(I+J) + 8(I+J), this rank's shard; "
                        "the pass reads 1 byte for every 2 it writes, and such a stream tops out below the copy rate "
                        "`peak` is measured with (mix_peak.mix_1r2w_gbs: same process, plain coalesced kernel)"}
    value = K / (ms * 1e-3)

    # ---------------- e2e: the public API with host buffers ----------------------
    e2e = None
    if not args.no_e2e:
        lib.scgbm_dev_free(ydev, device); ydev = None
        t0 = time.perf_counter()
        Yh, free_host, pinned = host_counts(lib, L, cfg, j0, j1, device, pinned=True)
        log(f"rank {rank}: host int32 Y ({Yh.nbytes / 1e9:.1f} GB, pinned={pinned}) ready in {time.perf_counter() - t0:.1f}s")
        R.barrier()
        with_se = cfg["what"] == "fit+se"
        t0 = time.perf_counter()
        out = gbm_sc(Yh, M, max_iter=args.e2e_max_iter, return_W=with_se, device=device, comm=R.comm, profile=2)
        fit_wall = time.perf_counter() - t0
        se_wall = None
        if with_se:
            t1 = time.perf_counter()
            out = get_se(out, device=device, comm=R.comm)
            se_wall = time.perf_counter() - t1
        wall = R.max(time.perf_counter() - t0)
        R.barrier()
        prof = out["_prof"]
        n_it = out["_n_iter"]
        # size-independent properties of the full-size fit (SURVEY section 4 iii), this rank's shard
        acc_obj = np.asarray(out["obj"])
        Uo, Vo = out["U"], out["V"]
        vtv = R.sum_array(Vo.T @ Vo)
        checks = {
            "obj_nondecreasing_from_3rd": bool(np.all(np.diff(acc_obj[2:]) >= -1e-9 * np.abs(acc_obj[2:-1]))) if len(acc_obj) > 3 else True,
            "mean_alpha_abs": float(abs(out["alpha"].mean())),
            "UtU_minus_I_max": float(np.abs(Uo.T @ Uo - np.eye(M)).max()),
            "VtV_minus_I_max": float(np.abs(vtv - np.eye(M)).max()),
            "D_descending": bool(np.all(np.diff(out["D"]) <= 0)),
            "U_first_row_nonneg": bool(np.all(Uo[0] >= 0)),
            "scores_eq_V_D_max": float(np.abs(out["scores"] - Vo * out["D"][None, :]).max()),
            "svd_unconverged_calls": out["_svd_unconverged"], "exact_rows": out["_exact_rows"],
            "underflow_reruns": out["_underflow_reruns"], "max_Y": out["_max_Y"],
        }
        if with_se:
            checks["se_scores_finite"] = bool(np.isfinite(out["se_scores"]).all())
        e2e = {"value": n_it / wall, "unit": "iters/s", "h2d_bytes_per_step": prof["h2d_bytes"] / n_it,
               "d2h_bytes_per_step": prof["d2h_bytes"] / n_it, "fit_seconds": fit_wall, "get_se_seconds": se_wall,
               "total_seconds": wall, "iterations": n_it,
               "accepted": int(len(out["obj"])), "device_ms": prof["total_ms"], "host_dtype": "int32",
               "pinned": bool(pinned), "objective_first_last": [float(out["obj"][0]), float(out["obj"][-1])],
               "fused_ms_total": round(prof["ms"]["fused"], 1), "checks": checks,
               "call": f"gbm_sc(Y, M={M}, max_iter={args.e2e_max_iter}, return_W={with_se})"
                       + (" then get_se(out)" if with_se else "") + " with reference defaults tol=1e-4, min_iter=30"}
        del Yh
        free_host()
    elif ydev is not None:
        lib.scgbm_dev_free(ydev, device)

    cb = None
    if rank == 0 and not args.no_cpu and world == 1:
        try:
            cb = cpu_baseline(cfg, 3, 1, args.cpu_sample_cells)
        except Exception as e:  # pragma: no cover
            log("cpu baseline failed:", e)
    if rank != 0:
        return None
    return {
        "metric": "gbm_sc_iters_per_sec", "value": value, "unit": "iters/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (eta from fp16x2-split tcgen05 MMAs, fp32 exp; f64 reductions and small dense)", "data": "synthetic",
        "config": config_dict(name, cfg, world, {
            "y_device_format": "u16 (s_Y=2)",
            "l2": "inputs (Y + residual planes, 6 B/entry) are far larger than the 126 MB L2; no flush needed"
                  if I * Jl * 6 > (1 << 30) else "shard smaller than 1 GB: L2 effects possible, see e2e",
            "svd": "warm block-Krylov, block M+12, adaptive depth, tol 1e-5 (irlba's own default); products on tcgen05",
            "svd_passes_per_step": svd_passes / K, "setup_and_initial_svd_seconds": setup_s}),
        "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(sum(launches.values())),
        "kernel_ms_per_step": {k: v / K for k, v in kms.items()},
        "kernel_ms_note": f"from {NB} extra fully instrumented steps after the timed region",
        "launches_per_step": {k: v / K for k, v in launches.items()},
    }


def run_se(name, cfg, args, R):
    """get.se alone (K8) at the c2 / c3 shape: seconds per call, lower is better."""
    from scgbm_b200 import _lib as L, gbm_sc, get_se
    lib = L.load()
    device = R.local_rank
    I, J, M = cfg["I"], cfg["J"], cfg["M"]
    if R.world != 1:
        raise SystemExit("the get.se configs are single-GPU")
    Yh, free_host, _ = host_counts(lib, L, cfg, 0, J, device)
    small = I * J * 8 <= (2 << 30)                      # W as an R matrix only where one could exist (c2: 160 MB)
    fit = gbm_sc(Yh, M, max_iter=40, return_W=small, keep_factors=not small, device=device)
    free_host()
    K, W = args.steps, max(args.warmup, 0)
    for _ in range(W):
        get_se(fit, device=device)
    sampler = ClockSampler(device); sampler.start()
    ts, prof = [], None
    for _ in range(K):
        t0 = time.perf_counter()
        out = get_se(fit, device=device, profile=True)
        ts.append(time.perf_counter() - t0)
        prof = out["_prof_se"]
    clocks = sampler.stop()
    sec = float(np.mean(ts))
    flops = 2.0 * I * J * M * (M + 1)                   # both Khatri-Rao Grams: N M (M+1) FMA
    cb = None
    if not args.no_cpu:
        try:
            cbf = cpu_baseline(cfg, 2, 1, args.cpu_sample_cells)
            cb = {"value": cbf["get_se_seconds_sample"] * cbf["extrapolation_factor"], "unit": "s", "cores": cbf["cores"],
                  "kind": "port", "sample": "oracle get_se_fast (vectorised R/uncertainty.R:14-35) on " + cbf["sample"],
                  "extrapolation_factor": cbf["extrapolation_factor"]}
        except Exception as e:  # pragma: no cover
            log("cpu baseline failed:", e)
    return {
        "metric": "get_se_seconds", "value": sec, "unit": "s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": sec * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(name, cfg, 1, {"input": "out$W (host doubles, uploaded)" if small else
                                             "kept factors (W re-evaluated on the device; no I x J host matrix)"}),
        "roofline": {"bound": "tensor", "achieved": flops / (prof["ms"]["se"] * 1e-3) / 1e12 if prof["ms"]["se"] > 0 else None,
                     "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                     "note": "K8 runs its Khatri-Rao Grams as fp64 FMAs on the CUDA cores (1e-9 parity with solve()); "
                             "achieved = 2 N M (M+1) flop / kernel time; no fp64 peak is measured for this pool"},
        "cpu_baseline": cb,
        "e2e": {"value": sec, "unit": "s", "h2d_bytes_per_step": prof["h2d_bytes"], "d2h_bytes_per_step": prof["d2h_bytes"],
                "device_ms": prof["total_ms"], "kernel_ms": prof["ms"]["se"]},
        "clocks": clocks, "gpu_launches": int(sum(prof["launches"].values())) * K,
    }


def _subset_index(starts, per):
    """1-based global column indices of the subsample"""
    return np.concatenate([np.arange(s, s + per) for s in starts]) + 1


def run_proj(name, cfg, args, R):
    """BASELINE configs[3]: gbm.sc(subset = 50 000 column indices) on 20k x 1M, cells sharded over the ranks.
    Every rank runs the identical subsample fit (bit-reproducible), projects its own cells (K9, one Poisson GLM per
    cell), and the only exchanges are the gene row sums and sum(exp(beta)) of R/fit.R:274."""
    from scgbm_b200 import _lib as L, gbm_proj_parallel
    from scgbm_b200.dist import shard_range
    lib = L.load()
    rank, world, device = R.rank, R.world, R.local_rank
    I, J, M, S = cfg["I"], cfg["J"], cfg["M"], cfg["subset"]
    j0, j1 = shard_range(J, rank, world)
    Jl = j1 - j0
    # the subsample: the first S/8 cells of each eighth of the matrix -- an explicit index vector (R/fit.R:221-223)
    # that any rank can regenerate by itself from the counter-based generator (no gather needed)
    blocks = 8
    per = S // blocks
    starts = [b * (J // blocks) for b in range(blocks)]
    t0 = time.perf_counter()
    Y_sub = None
    if world > 1:
        Y_sub = np.asfortranarray(np.hstack([host_counts(lib, L, cfg, s0, s0 + per, device)[0] for s0 in starts]))
    Yh, free_host, pinned = host_counts(lib, L, cfg, j0, j1, device, pinned=True)
    log(f"rank {rank}: shard {Yh.shape} (+ subsample) on the host in {time.perf_counter() - t0:.1f}s")
    K, W = max(1, min(args.steps, 3)), 0
    sampler = ClockSampler(device); sampler.start()
    walls, out = [], None
    for _ in range(K):
        R.barrier()
        t0 = time.perf_counter()
        if world > 1:
            out = gbm_proj_parallel(Yh, M, device=device, comm=R.comm, subsample_Y=Y_sub, profile=True)
        else:
            out = gbm_proj_parallel(Yh, M, subsample=_subset_index(starts, per), device=device, profile=True)
        walls.append(R.max(time.perf_counter() - t0))
    clocks = sampler.stop()
    wall = float(np.mean(walls))
    pp = out["_prof_proj"]
    # K9's device time: every kernel class of the scgbm_project call except the upload / pack of the counts (the
    # tensor-core path spends it in the fused pass's product rider, class "fused"; the exact kernel in class "proj")
    k9_ms = float(sum(v for k, v in pp["ms"].items() if k != "pack"))
    proj_ms = R.max(k9_ms)
    fit_ms = out["_prof"]["total_ms"]
    Ik = len(out["_ixs"])
    sweeps = float(np.mean(out["_glm_iters"]))
    flops = Jl * Ik * ((M + 1) * (M + 2) + 4 * (M + 1)) * sweeps * 2.0          # BASELINE.md section 4, FMA = 2 flop
    fails = int(out["_fail"].sum())
    nbytes = float(Yh.nbytes + (Y_sub.nbytes if Y_sub is not None else 0))
    # the tensor-core path's rider launches in three views: each reads the shard's counts once (u16 unless the shard's
    # maximum allowed u8: bytes quoted at 2 B, an upper bound), evaluates one exp per entry, and writes nothing I x J
    tc = out.get("_glm_tc") or {}
    n_rider = int(tc.get("sweeps", 0)) * int(math.ceil(((M + 1) * (M + 2) / 2 + 1) / 32.0))
    rider_ms = float(pp["ms"].get("fused", 0.0))
    views = None
    if out["_glm_method"] == "tcgen05" and n_rider > 0 and rider_ms > 0:
        peak_gbs, _src = measured_peak()
        entries = float(n_rider) * Jl * Ik
        views = {"rider_launches": n_rider, "rider_ms": rider_ms,
                 "hbm": {"achieved_gbs": entries * 2.0 / (rider_ms * 1e-3) / 1e9, "peak_gbs": peak_gbs,
                         "frac": entries * 2.0 / (rider_ms * 1e-3) / 1e9 / peak_gbs if peak_gbs else None},
                 "exp_per_s": entries / (rider_ms * 1e-3),
                 "bound": "epilogue issue / SFU, not HBM: every 32-column block of the Khatri-Rao operand re-evaluates the eta tile "
                          "and the exp (DESIGN.md section 10, item 4: one launch per sweep would cut this 8x)"}
    free_host()
    if rank != 0:
        return None
    return {
        "metric": "gbm_sc_subset_cells_per_sec", "value": J / (proj_ms * 1e-3), "unit": "cells/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": proj_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (eta and products from fp16x2-split tcgen05 MMAs, fp32 exp; fp64 normal equations)" if out["_glm_method"] == "tcgen05" else "f64",
        "data": "synthetic",
        "config": config_dict(name, cfg, world, {"subset": f"{S} explicit column indices ({blocks} blocks of {per})",
                                                 "genes_kept": int(Ik), "irls_sweeps_mean": sweeps, "glm_failures": fails}),
        "roofline": {"bound": "tensor", "achieved": flops / (k9_ms * 1e-3) / 1e12, "peak": None, "unit": "TFLOP/s",
                     "frac": None, "traffic": None, "glm_method": out["_glm_method"], "tc": out["_glm_tc"],
                     "kernel_ms": {k: v for k, v in pp["ms"].items() if v > 0}, "rider_views": views,
                     "note": "K9, rank 0: algorithmic IRLS flops (BASELINE.md section 4, at the mean number of solves per cell) over "
                             "the device time of the projection kernels.  tcgen05 method (proj_tc.cu): all cells in lock step, "
                             "gradient and Hessian of every Newton sweep from ceil((T+1)/32) launches of fused_tc_kernel's product "
                             "rider (eta tile and both products on the tensor cores), the last ~2 % of slow cells by the exact fp64 "
                             "kernel; exact method (proj_irls_kernel): fp64 FMAs on the CUDA cores, fastglm's own iteration"},
        "cpu_baseline": None,
        "e2e": {"value": J / wall, "unit": "cells/s", "seconds": wall, "subsample_fit_device_ms": fit_ms, "projection_ms": proj_ms,
                "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": float(Jl * (2 * M + 1) * 8),
                "call": f"gbm_sc(Y, M={M}, subset=idx): subsample fit on {Ik}x{S} + projection of {J} cells; host int32 buffers",
                "checks": {"mean_alpha_abs": float(abs(out["alpha"].mean())), "scores_finite": bool(np.isfinite(out["scores"]).all())}},
        "clocks": clocks, "gpu_launches": int(sum(pp["launches"].values()) + sum(out["_prof"]["launches"].values())),
    }


# ----------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5p", choices=sorted(CONFIGS))
    ap.add_argument("--genes", type=int, default=int(os.environ.get("SCGBM_BENCH_GENES", 0)))     # overrides (experiments)
    ap.add_argument("--cells", type=int, default=int(os.environ.get("SCGBM_BENCH_CELLS", 0)))
    ap.add_argument("--factors", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-max-iter", type=int, default=100)
    ap.add_argument("--cpu-sample-cells", type=int, default=2000)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    name = args.config
    if args.genes or args.cells or args.factors:
        if args.genes: cfg["I"] = args.genes
        if args.cells: cfg["J"] = args.cells
        if args.factors: cfg["M"] = cfg["d"] = args.factors
        name += " (shape overridden)"
    R = Ranks()
    K, W = args.steps, max(args.warmup, 0)

    if args.impl == "reference":
        if R.rank != 0:
            return
        if cfg["what"] == "proj":
            print(json.dumps({"impl": "reference", "unavailable": "the oracle's per-cell GLM loop at 1M cells does not fit the time "
                              "box; parity of the projection is tests/test_gpu_fit.py::test_projection_matches_oracle"}))
            return
        cb = cpu_baseline(cfg, K, W, args.cpu_sample_cells)
        if cfg["what"] == "se":
            val, metric, unit, hib = cb["get_se_seconds_sample"] * cb["extrapolation_factor"], "get_se_seconds", "s", False
        else:
            val, metric, unit, hib = cb["value"], "gbm_sc_iters_per_sec", "iters/s", True
        line = {"impl": "reference", "metric": metric, "value": val, "unit": unit,
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1000.0 / val if hib else val * 1e3,
                "higher_is_better": hib, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(name, cfg, args.gpus), "cpu_baseline": cb,
                # the slice the CPU arm really ran and the factor its value was scaled by: a projection, not a measurement
                "sample_shape": cb["sample_shape"], "extrapolation_factor": cb["extrapolation_factor"],
                "extrapolated": cb["extrapolation_factor"] > 1.0001, "blas": cb["blas"], "cores": cb["cores"],
                "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    R.init(R.local_rank)
    try:
        if cfg["what"] == "se":
            line = run_se(name, cfg, args, R)
        elif cfg["what"] == "proj":
            line = run_proj(name, cfg, args, R)
        else:
            line = run_fit(name, cfg, args, R)
    finally:
        R.close()
    if R.rank == 0 and line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
