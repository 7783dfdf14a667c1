"""CPU tests of the N>1 host logic on a world_size-2 gloo group: the flat-buffer gradient all-reduce (also when the
ranks disagree about which parameters received a gradient), the rank-0 broadcast of parameters and buffers that
DDP performs at construction / before every forward, and the patch dealing of the
sharded sliding-window inference."""
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
    from unet_bssfp_b200.train_step import GradAllReducer, broadcast_module_state
    ok = True
    # ---- (1) all-reduce with a frozen slot
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    frozen = net[0].bias
    frozen.requires_grad_(False)                      # frozen / unused slots are skipped, as DDP does
    x = torch.full((5, 4), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    scale = GradAllReducer(net)()
    got = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    want = [sum(g[i] for g in gathered) / world for i in range(len(local))]
    ok = ok and scale == 1.0 and all(torch.allclose(a, b, rtol=1e-6, atol=1e-7) for a, b in zip(got, want)) and frozen.grad is None
    # ... and with the scale left to the optimizer (FusedAdamW.grad_scale): the buffer holds the SUM
    for p in net.parameters():
        p.grad = None
    net(x).sum().backward()
    scale = GradAllReducer(net)(apply_scale=False)
    got = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    ok = ok and scale == 1.0 / world and all(torch.allclose(a * scale, b, rtol=1e-6, atol=1e-7) for a, b in zip(got, want))
    # ---- (2) the ranks disagree: rank 1 never touches the second branch (find_unused_parameters semantics)
    torch.manual_seed(0)
    two = torch.nn.ModuleDict({"a": torch.nn.Linear(4, 2), "b": torch.nn.Linear(4, 2)})
    out = two["a"](x).sum() + (two["b"](x).sum() if rank == 0 else 0.0)
    out.backward()
    gb0 = two["b"].weight.grad.clone() if rank == 0 else torch.zeros_like(two["b"].weight)
    GradAllReducer(two, check_unused=True)()          # must not hang; missing slots count as zeros
    gb = [None] * world
    dist.all_gather_object(gb, gb0)
    ok = ok and two["b"].weight.grad is not None and torch.allclose(two["b"].weight.grad, sum(gb) / world)
    # ---- (3) rank-0 broadcast of parameters and buffers
    torch.manual_seed(100 + rank)                     # different initial state per rank
    bn = torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.BatchNorm1d(3))
    bn[1].running_mean.add_(float(rank))
    bn[1].num_batches_tracked.add_(rank * 5)
    state0 = [None] * world
    dist.all_gather_object(state0, {k: v.clone() for k, v in bn.state_dict().items()})
    ptrs = [t.data_ptr() for t in bn.state_dict().values()]
    broadcast_module_state(bn)
    ok = ok and all(torch.equal(v, state0[0][k]) for k, v in bn.state_dict().items())
    ok = ok and ptrs == [t.data_ptr() for t in bn.state_dict().values()]          # in place
    bn[1].running_var.mul_(float(rank + 2)); bn[0].weight.data.add_(float(rank))
    w_before = bn[0].weight.detach().clone()
    broadcast_module_state(bn, buffers_only=True)     # buffers follow rank 0, parameters are left alone
    rv = [None] * world
    dist.all_gather_object(rv, bn[1].running_var.clone())
    ok = ok and torch.equal(rv[0], rv[1]) and torch.equal(bn[0].weight.detach(), w_before)
    # ---- (4) patch dealing of the sharded inference
    from unet_bssfp_b200.inference import grid_locations, shard_indices
    origins = grid_locations((160, 192, 160), (64, 64, 64))
    mine = shard_indices(len(origins), rank, world)
    allm = [None] * world
    dist.all_gather_object(allm, mine)
    ok = ok and len(origins) == 27 and sorted(sum(allm, [])) == list(range(27)) and mine == sorted(mine)
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
