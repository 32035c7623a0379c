"""Rendezvous helper for the multi-process tests: hands every rank the NCCL unique id of libscgbm over an
initialised torch.distributed (gloo) group.  Lives with the tests (and, duplicated, in bench.py): the product
package scgbm_b200/ imports no torch."""


def torch_broadcast_bytes(nbytes: int = 128, src: int = 0):
    """broadcast_bytes callable for scgbm_b200.dist.init_comm on top of torch.distributed."""
    import torch
    import torch.distributed as dist

    def _bc(payload):
        t = torch.zeros(nbytes, dtype=torch.uint8)
        if payload is not None:
            t = torch.tensor(list(payload), dtype=torch.uint8)
        dist.broadcast(t, src=src)
        return bytes(t.tolist())
    return _bc
