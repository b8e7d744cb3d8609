"""CPU, world_size 2, gloo: the N>1 host path -- batch sharding + the single output gather.
The per-rank forward is the torch port (test infrastructure) standing in for the CUDA module."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lpsr_b200
        from oracle import lpsr_torch_port as tp
        W = tp.to_torch_weights(dict(np.load(os.path.join(GOLDEN, "weights_best_model.npz"))))
        x = torch.rand(batch, 3, 8, 16, generator=torch.Generator().manual_seed(77))
        fwd = lambda t: tp.lpsr_forward(t, W) if t.shape[0] else torch.empty(0, 1, 8, 16)
        y = lpsr_b200.forward_sharded(fwd, x, gather=True)
        lo, hi = lpsr_b200.shard_bounds(batch, world, rank)
        full = tp.lpsr_forward(x, W)
        ok = y.shape == full.shape and torch.allclose(y, full, atol=1e-6) and torch.equal(y[lo:hi], fwd(x[lo:hi]))
        q.put((rank, bool(ok), tuple(y.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5, 1])
def test_sharded_forward_gathers_full_batch_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (batch, 1, 8, 16) for _, _, shape in res)


# ---- GPU: the same sharded forward with the CUDA module on every rank, NCCL gather over NVLink ----------------------------
def _gpu_worker(rank, world, port, batch, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lpsr_b200
        W = dict(np.load(os.path.join(GOLDEN, "weights_best_model.npz")))
        m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="fp16")
        m.load_live_weights(W)
        m = m.to(f"cuda:{rank}").eval()
        x = torch.rand(batch, 3, 32, 96, generator=torch.Generator().manual_seed(78)).to(f"cuda:{rank}")
        y = lpsr_b200.forward_sharded(m, x, gather=True)          # this rank's shard on its GPU + ONE all-gather
        full = m(x)                                               # the whole batch on this GPU alone
        torch.cuda.synchronize()
        q.put((rank, bool(torch.equal(y, full)), tuple(y.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [8, 5])
def test_sharded_forward_cuda_module_nccl(batch):
    """SURVEY 8e on real GPUs: batch-sharded forward + the one NCCL all-gather equals the single-GPU result bit for bit
    (a crop's result does not depend on its batch or its GPU).  Skipped on a one-GPU box."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (batch, 1, 32, 96) for _, _, shape in res)
