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
    from scgbm_b200.dist import shard_range, torch_broadcast_bytes
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
