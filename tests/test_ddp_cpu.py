"""CPU test of the N>1 host logic: the flat-buffer gradient all-reduce on a world_size-2 gloo group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet_bssfp_b200.train_step import GradAllReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    frozen = net[0].bias
    frozen.requires_grad_(False)                      # frozen / unused slots are skipped, as DDP does
    x = torch.full((5, 4), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    GradAllReducer(net)()
    got = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    want = [sum(g[i] for g in gathered) / world for i in range(len(local))]
    ok = all(torch.allclose(a, b, rtol=1e-6, atol=1e-7) for a, b in zip(got, want)) and frozen.grad is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_grad_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
