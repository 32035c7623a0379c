"""world_size-2 gloo test of the N>1 host logic: column sharding, unique-id broadcast
plumbing, and that the exchanges the multi-GPU path performs (gene-side
row sums, objective scalars, max) reproduce the single-process oracle quantities."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from scgbm_b200.dist import shard_range
    from tests.dist_util import torch_broadcast_bytes
    from oracle.r_rng import readme_demo_Y
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (b) unique-id style broadcast: every rank ends up with rank 0's 128 bytes
        payload = bytes(range(128)) if rank == 0 else None
        got = torch_broadcast_bytes()(payload)
        assert got == bytes(range(128))
        # (a) shards tile the columns
        Y = readme_demo_Y().astype(np.float64)
        I, J = Y.shape
        j0, j1 = shard_range(J, rank, world)
        Yl = Y[:, j0:j1]
        # (c) the per-iteration exchanges, R/fit.R:127-151 with X = 0
        rs = torch.tensor(Yl.sum(1)); dist.all_reduce(rs)                  # rowSums(Y): gene-side all-reduce
        beta = np.log(Yl.sum(0))                                           # colSums are local
        S = torch.tensor(np.exp(beta)[None, :].repeat(I, 0).sum(1)); dist.all_reduce(S)
        alpha = np.log(rs.numpy()) - np.log(S.numpy())
        am = alpha.mean(); alpha -= am; beta += am
        W = np.exp(alpha[:, None] + beta[None, :])
        mY = torch.tensor([Yl.max()]); dist.all_reduce(mY, op=dist.ReduceOp.MAX)
        W = np.minimum(W, mY.item())
        nz = Yl != 0
        part = torch.tensor([(Yl[nz] * np.log(W[nz])).sum(), W.sum()]); dist.all_reduce(part)
        wmax = torch.tensor([W.max()]); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        q.put((rank, j0, j1, float(part[0] - part[1]), float(wmax), alpha.copy(), beta.copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchanges_match_single_process():
    import torch.multiprocessing as mp
    from oracle.r_rng import readme_demo_Y
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(60); assert p.exitcode == 0
    # single-process restatement
    Y = readme_demo_Y().astype(np.float64)
    beta = np.log(Y.sum(0)); alpha = np.log(Y.sum(1)) - np.log(np.exp(beta).sum())
    am = alpha.mean(); alpha -= am; beta += am
    W = np.minimum(np.exp(alpha[:, None] + beta[None, :]), Y.max())
    LL = (Y[Y != 0] * np.log(W[Y != 0])).sum() - W.sum()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == Y.shape[1]
    for r in res:
        assert abs(r[3] - LL) < 1e-9 * abs(LL)
        assert abs(r[4] - W.max()) < 1e-12
        np.testing.assert_allclose(r[5], alpha, atol=1e-12)
    np.testing.assert_allclose(np.concatenate([res[0][6], res[1][6]]), beta, atol=1e-12)


@pytest.mark.parametrize("J,n", [(10, 3), (500, 8), (7, 8), (1000000, 8), (1, 1)])
def test_shard_range_tiles(J, n):
    from scgbm_b200.dist import shard_range
    edges = [shard_range(J, r, n) for r in range(n)]
    assert edges[0][0] == 0 and edges[-1][1] == J
    for a, b in zip(edges, edges[1:]):
        assert a[1] == b[0]
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------
# Cell-sharded gbm.sc(subset=) (BASELINE configs[3]): the host logic of
# scgbm_b200/projection.py -- subsample handling, row filter, the two all-reduces and the re-centring of
# R/fit.R:272-277 -- on two gloo ranks.  The two GPU entry points it calls are stood in for by the oracle
# (tests may do that; the product never does), so this runs without a device.
# ---------------------------------------------------------------------------------------------
def _proj_worker(rank, world, port, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import scgbm_b200.api as api
    from scgbm_b200 import _lib as L
    from scgbm_b200.dist import shard_range
    from scgbm_b200.projection import gbm_proj_parallel
    from oracle import scgbm_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Y = O.sim_data(60, 90, 2, seed=4)["Y"].astype(np.int32)
        Y[5, :] = 0; Y[5, 1] = 3                       # a gene that the > 5 row filter (R/fit.R:225) drops
        jxs = np.arange(0, 90, 2)
        j0, j1 = shard_range(Y.shape[1], rank, world)

        def fake_fit(Ysub, M, **kw):                   # stands in for scgbm_fit
            return O.gbm_sc(np.asarray(Ysub, dtype=np.float64), M, tol=kw.get("tol", 1e-4), max_iter=kw.get("max_iter", 100))

        def fake_project(a_ref, o_ref):                # stands in for scgbm_project (R/fit.R:248-264)
            a, o = a_ref._obj, o_ref._obj
            assert a.y_dtype == L.I32 and not a.y_on_device
            J, Ik, M = a.J, a.Ik, a.M
            Yl = np.ctypeslib.as_array(C.cast(a.Y, C.POINTER(C.c_int32)), shape=(J, a.ldY)).T.astype(np.float64)
            ixs = np.ctypeslib.as_array(C.cast(a.ixs, C.POINTER(C.c_int64)), shape=(Ik,))
            U = np.ctypeslib.as_array(C.cast(a.U, C.POINTER(C.c_double)), shape=(M, Ik)).T
            alpha = np.ctypeslib.as_array(C.cast(a.alpha, C.POINTER(C.c_double)), shape=(Ik,))
            beta = np.ctypeslib.as_array(C.cast(o.beta, C.POINTER(C.c_double)), shape=(J,))
            sc = np.ctypeslib.as_array(C.cast(o.scores, C.POINTER(C.c_double)), shape=(M, J))
            se = np.ctypeslib.as_array(C.cast(o.se_scores, C.POINTER(C.c_double)), shape=(M, J))
            if o.rowsum_full:                          # rowSums(Y) of all genes over this rank's cells (R/fit.R:217)
                rsf = np.ctypeslib.as_array(C.cast(o.rowsum_full, C.POINTER(C.c_double)), shape=(a.I_full,))
                rsf[:] = Yl.sum(axis=1)
            X = np.hstack([np.ones((Ik, 1)), U])
            for j in range(J):
                coef, s, _, _ = O.poisson_glm_irls(X, Yl[ixs, j], alpha)
                beta[j] = coef[0]; sc[:, j] = coef[1:]; se[:, j] = s[1:]
            return 0

        class GlooComm:                                # the two collectives Comm offers, over gloo
            handle = None

            def allreduce(self, arr, op="sum"):
                t = torch.tensor(np.ascontiguousarray(arr, dtype=np.float64))
                dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
                return t.numpy()

        api.gbm_sc = fake_fit
        L.load().scgbm_project = fake_project
        out = gbm_proj_parallel(np.asfortranarray(Y[:, j0:j1]), 2, comm=GlooComm(),
                                subsample_Y=np.asfortranarray(Y[:, jxs]), max_iter=15)
        q.put((rank, out["alpha"], out["beta"], out["scores"], out["se_scores"], out["U"], out["_ixs"]))
    finally:
        dist.destroy_process_group()


def test_sharded_projection_host_logic():
    import torch.multiprocessing as mp
    from oracle import scgbm_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_proj_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=240) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(60); assert p.exitcode == 0
    Y = O.sim_data(60, 90, 2, seed=4)["Y"].astype(np.int32)
    Y[5, :] = 0; Y[5, 1] = 3
    ref = O.gbm_proj_parallel(Y.astype(np.float64), 2, subsample=np.arange(0, 90, 2) + 1, max_iter=15)
    assert 5 not in ref["_ixs"]
    for r in res:                                       # replicated on every rank
        np.testing.assert_array_equal(r[6], ref["_ixs"])
        np.testing.assert_allclose(r[1], ref["alpha"], atol=1e-10)
        np.testing.assert_allclose(r[5], ref["U"], atol=1e-12)
    np.testing.assert_allclose(np.concatenate([res[0][2], res[1][2]]), ref["beta"], atol=1e-10)
    np.testing.assert_allclose(np.vstack([res[0][3], res[1][3]]), ref["scores"], atol=1e-10)
    np.testing.assert_allclose(np.vstack([res[0][4], res[1][4]]), ref["se_scores"], rtol=1e-9)


# ---------------------------------------------------------------------------------------------
# Batch levels must be GLOBAL when the cells are sharded: a shard that lacks a level must not
# re-number the levels it has.  Host logic of scgbm_b200/api.py::_batch_codes on two gloo ranks.
# ---------------------------------------------------------------------------------------------
def _levels_worker(rank, world, port, q):
    import pickle
    import torch
    import torch.distributed as dist
    from scgbm_b200.api import _batch_codes, RError
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class GlooComm:
            nranks = world

            def allgather_bytes(self, payload):
                out = [None] * world
                dist.all_gather_object(out, payload)
                return out
        labels = np.array(["A", "B", "A", "B"]) if rank == 0 else np.array(["B", "B", "B"])      # rank 1 lacks level A
        codes, nb = _batch_codes(labels, len(labels), comm=GlooComm())
        num = np.array([3, 7, 3]) if rank == 0 else np.array([5, 5, 7])
        ncodes, nnb = _batch_codes(num, 3, comm=GlooComm())
        none_codes, none_nb = _batch_codes(None, 3, comm=GlooComm())
        try:                        # batch given on one rank only: refused on both
            _batch_codes(labels if rank == 0 else None, len(labels) if rank == 0 else 3, comm=GlooComm())
            mixed = "accepted"
        except RError as e:
            mixed = str(e)
        q.put((rank, codes.tolist(), nb, ncodes.tolist(), nnb, none_codes, none_nb, mixed))
    finally:
        dist.destroy_process_group()


def test_batch_levels_are_agreed_across_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_levels_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(60); assert p.exitcode == 0
    assert res[0][1] == [1, 2, 1, 2] and res[1][1] == [2, 2, 2]        # B is level 2 on BOTH ranks
    assert res[0][2] == res[1][2] == 2
    assert res[0][3] == [1, 3, 1] and res[1][3] == [2, 2, 3] and res[0][4] == res[1][4] == 3
    assert res[0][5] is None and res[1][5] is None and res[0][6] == res[1][6] == 1
    assert "every rank or on none" in res[0][7] and "every rank or on none" in res[1][7]
