"""N-GPU parity: cells sharded over NCCL ranks give the same fit as one GPU.
Skipped unless the box has at least two GPUs (`gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, kw):
    import torch.distributed as dist
    from scgbm_b200 import gbm_sc
    from scgbm_b200.dist import shard_range, init_comm, torch_broadcast_bytes
    from oracle.scgbm_oracle import sim_data
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Y = sim_data(700, 1800, 8, seed=5, nb_theta=1.0)["Y"]
        j0, j1 = shard_range(Y.shape[1], rank, world)
        comm = init_comm(rank, world, rank, torch_broadcast_bytes())
        out = gbm_sc(np.asfortranarray(Y[:, j0:j1]), 8, device=rank, comm=comm, diagnostics=True, return_W=False, **kw)
        comm.destroy()
        q.put((rank, j0, j1, out["_LL"], out["_accepted"], out["alpha"], out["beta"], out["U"], out["V"], out["D"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kw", [dict(), dict(infer_beta=True, max_iter=25)])
def test_two_gpus_match_one(kw):
    from scgbm_b200 import _lib, gbm_sc
    from oracle.scgbm_oracle import sim_data
    from tests.util import principal_angle
    if _lib.load().scgbm_device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q, kw)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(60); assert p.exitcode == 0
    Y = sim_data(700, 1800, 8, seed=5, nb_theta=1.0)["Y"]
    one = gbm_sc(Y, 8, diagnostics=True, return_W=False, **kw)
    for r in res:       # replicated quantities are identical on every rank
        n = min(len(r[3]), len(one["_LL"]))
        np.testing.assert_array_equal(r[4][:n], one["_accepted"][:n])
        np.testing.assert_allclose(r[3][:n], one["_LL"][:n], rtol=1e-6)
        np.testing.assert_allclose(r[5], one["alpha"], atol=1e-5)
        np.testing.assert_allclose(r[9], one["D"], rtol=1e-5)
        assert principal_angle(r[7], one["U"]) < 1e-4
    np.testing.assert_array_equal(res[0][7], res[1][7])
    beta = np.concatenate([res[0][6], res[1][6]])
    V = np.vstack([res[0][8], res[1][8]])
    np.testing.assert_allclose(beta, one["beta"], atol=1e-5)
    assert principal_angle(V, one["V"]) < 1e-4


def _proj_worker(rank, world, port, q):
    import torch.distributed as dist
    from scgbm_b200 import gbm_proj_parallel
    from scgbm_b200.dist import shard_range, init_comm, torch_broadcast_bytes
    from oracle.scgbm_oracle import sim_data
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Y = sim_data(300, 900, 4, seed=9)["Y"]
        jxs = np.arange(0, 900, 3)                   # the subsample every rank is handed
        j0, j1 = shard_range(Y.shape[1], rank, world)
        comm = init_comm(rank, world, rank, torch_broadcast_bytes())
        out = gbm_proj_parallel(np.asfortranarray(Y[:, j0:j1]), 4, device=rank, comm=comm,
                                subsample_Y=np.asfortranarray(Y[:, jxs]), max_iter=40)
        comm.destroy()
        q.put((rank, out["alpha"], out["beta"], out["scores"], out["se_scores"], out["U"]))
    finally:
        dist.destroy_process_group()


def test_sharded_projection_matches_one_gpu():
    """BASELINE configs[3]: gbm.sc(subset=) with the cells sharded; R/fit.R:213-285."""
    from scgbm_b200 import _lib, gbm_proj_parallel
    from oracle.scgbm_oracle import sim_data
    if _lib.load().scgbm_device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_proj_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(60); assert p.exitcode == 0
    Y = sim_data(300, 900, 4, seed=9)["Y"]
    one = gbm_proj_parallel(Y, 4, subsample=np.arange(0, 900, 3) + 1, max_iter=40)
    np.testing.assert_array_equal(res[0][5], res[1][5])          # the subsample fit is replicated bit for bit
    np.testing.assert_allclose(res[0][5], one["U"], atol=1e-12)
    for r in res:
        np.testing.assert_allclose(r[1], one["alpha"], atol=1e-10)
    np.testing.assert_allclose(np.concatenate([res[0][2], res[1][2]]), one["beta"], atol=1e-10)
    np.testing.assert_allclose(np.vstack([res[0][3], res[1][3]]), one["scores"], atol=1e-10)
    np.testing.assert_allclose(np.vstack([res[0][4], res[1][4]]), one["se_scores"], rtol=1e-9)
