"""Data-parallel correctness on hardware (VERDICT r1 missing #3 / #4), two processes on a gloo group sharing the
box's GPU (NCCL refuses two ranks on one device; the kernels, autograd hooks and reducers exercised are the same):

* the drop-in modules under the REAL ``torch.nn.parallel.DistributedDataParallel(find_unused_parameters=True)`` --
  what Lightning's ``strategy='ddp_find_unused_parameters_true'`` builds (ref:src/train.py:30-32) -- with the
  reference's phase toggling (ref:src/model.py:264,274): construction broadcasts rank 0's state,
  and after ``backward`` every rank holds the MEAN of the per-shard gradients of the same weights (C1 / C2);
* ``GanTrainer`` (the plain-``torch.distributed`` harness bench.py uses): replicas stay identical, and one step equals
  a single-process emulation "each rank == single GPU on its shard, gradients averaged" bit for bit;
* the rank-sharded sliding-window inference equals the single-process result bit for bit.
"""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _no_dropout(*mods):
    for mod in mods:
        for m in mod.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0


class _Both(torch.nn.Module):
    """The two networks behind one forward, as the LightningModule DDP wraps (training_step runs inside forward)."""

    def __init__(self, gen, discr):
        super().__init__()
        import unet_bssfp_b200 as ub
        self.gen, self.discr = gen, discr
        self.l1, self.bce = ub.L1Loss(), ub.BCEWithLogitsLoss()

    def forward(self, x, y, phase):
        if phase == "g":                                   # ref:src/model.py:170-181
            y_hat = self.gen(x)
            logits = self.discr(x, y_hat)
            return self.bce(logits, torch.ones_like(logits)) + self.l1(y_hat, y) / 2 * 1e2
        with torch.no_grad():                              # ref:src/model.py:183-193
            y_hat = self.gen(x)
        lh, lr = self.discr(x, y_hat), self.discr(x, y)
        return (self.bce(lr, torch.ones_like(lr)) + self.bce(lh, torch.zeros_like(lh))) / 2


def _toggle(both, phase):
    for p in both.gen.parameters():
        p.requires_grad_(phase == "g")
    for p in both.discr.parameters():
        p.requires_grad_(phase == "d")


def _grads(mod):
    return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in mod.named_parameters()}


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        import unet_bssfp_b200 as ub
        from unet_bssfp_b200.train_step import GanTrainer
        res = {}
        torch.manual_seed(100 + rank)                                  # every rank starts from DIFFERENT weights
        both = _Both(ub.Generator("bssfp"), ub.Discriminator("bssfp")).to(dev)
        _no_dropout(both)
        torch.manual_seed(7 + rank)                                    # ... and draws its own shard
        x = torch.rand(2, 24, 32, 32, 32, device=dev)
        y = torch.rand(2, 6, 32, 32, 32, device=dev)

        # ---------------- real DistributedDataParallel ----------------
        ddp = torch.nn.parallel.DistributedDataParallel(both, find_unused_parameters=True, broadcast_buffers=True)
        sums = [None] * world
        dist.all_gather_object(sums, float(sum(p.double().sum() for p in both.parameters())))
        res["ddp_broadcast_at_construction"] = sums[0] == sums[1]
        plain = copy.deepcopy(both)                                    # same weights, no DDP: the per-shard truth
        ok = True
        for phase in ("g", "d"):
            _toggle(both, phase); _toggle(plain, phase)
            for m in (both, plain):
                for p in m.parameters():
                    p.grad = None
            ddp(x, y, phase).backward()
            plain(x, y, phase).backward()
            mine = _grads(plain)
            shards = [None] * world
            dist.all_gather_object(shards, {k: (None if v is None else v.cpu()) for k, v in mine.items()})
            got = _grads(both)
            live = [k for k, v in mine.items() if v is not None]
            ok = ok and len(live) > 10 and all((got[k] is None) == (mine[k] is None) for k in mine)
            for k in live:
                want = sum(s[k] for s in shards) / world
                ok = ok and torch.allclose(got[k].cpu(), want, rtol=1e-5, atol=1e-7 * float(want.abs().max()) + 1e-12)
            # the other network and the unused input heads got nothing
            other = "discr." if phase == "g" else "gen."
            ok = ok and all(v is None for k, v in got.items() if k.startswith(other))
        res["ddp_grads_are_shard_means"] = ok
        del ddp, plain

        # ---------------- GanTrainer: broadcast + flat all-reduce + AdamW ----------------
        torch.manual_seed(200 + rank)
        g, d = ub.Generator("bssfp").to(dev), ub.Discriminator("bssfp").to(dev)
        _no_dropout(g, d)
        tr = GanTrainer(g, d)                                          # broadcasts rank 0's parameters and buffers
        g0, d0 = copy.deepcopy(g), copy.deepcopy(d)                    # state before the step (identical on all ranks)
        tr.step(x, y)
        torch.cuda.synchronize()
        gp = [None] * world
        dist.all_gather_object(gp, [p.detach().cpu() for p in g.parameters()] + [p.detach().cpu() for p in d.parameters()])
        res["trainer_replicas_identical"] = all(torch.equal(a, b) for a, b in zip(gp[0], gp[1]))
        # single-process emulation of the G phase: each shard on the same weights, gradients summed, 1/world in AdamW
        shards = [None] * world
        dist.all_gather_object(shards, (x.cpu(), y.cpu()))
        if rank == 0:
            opt = ub.FusedAdamW(g0.parameters(), lr=1e-3)
            total = None
            for xs, ys in shards:
                gg, dd = copy.deepcopy(g0), copy.deepcopy(d0)          # BatchNorm buffers of every shard start equal
                for p, q_ in zip(gg.parameters(), g0.parameters()):
                    assert p.data_ptr() != q_.data_ptr()
                t1 = GanTrainer.__new__(GanTrainer)
                t1.gen, t1.discr, t1.l1, t1.bce = gg, dd, ub.L1Loss(), ub.BCEWithLogitsLoss()
                for p in dd.parameters():
                    p.requires_grad_(False)
                loss, _ = t1.gen_loss(xs.to(dev), ys.to(dev))
                loss.backward()
                gs = [p.grad for p in gg.parameters()]
                total = gs if total is None else [a if b is None else (b if a is None else a + b) for a, b in zip(total, gs)]
            for p, gsum in zip(g0.parameters(), total):
                p.grad = gsum
            opt.grad_scale = 1.0 / world
            opt.step()
            res["trainer_step_equals_emulation"] = all(torch.equal(a.detach(), b.detach())
                                                       for a, b in zip(g.parameters(), g0.parameters()))
        else:
            res["trainer_step_equals_emulation"] = True

        # ---------------- rank-sharded sliding-window inference ----------------
        # every rank must hold the same weights AND buffers: the step above left each rank's BatchNorm running
        # statistics updated from its own shard (rank 0's are broadcast at the start of the next step, as DDP does)
        from unet_bssfp_b200.train_step import broadcast_module_state
        broadcast_module_state(g, buffers_only=True)
        g.eval()
        torch.manual_seed(5)                                           # the same volume on every rank
        vol = torch.rand((24, 48, 80, 40), device=dev)                # 2 x 3 x 2 = 12 patches of 32^3, overlapping at the ends
        # one patch per launch: the same launch geometry whoever runs the patch -> bit-identical aggregation
        whole = ub.inference.predict_volume(g, vol, patch=32, batch=1, shard=False)
        sharded = ub.inference.predict_volume(g, vol, patch=32, batch=1)           # shards: dist is initialised
        again = ub.inference.predict_volume(g, vol, patch=32, batch=1, shard=False)
        res["inference_is_deterministic"] = bool(torch.equal(whole, again))
        same = bool(torch.equal(whole, sharded))
        res["sharded_inference_bit_identical"] = same if same else (
            f"max |diff| {(whole - sharded).abs().max().item():.3e}, {int((whole != sharded).sum())} of {whole.numel()} differ")
        # batched: the split-K / depth-segment geometry depends on the batch size, so fp32 sums associate differently
        # (bf16 noise level); a wrong owner or a lost patch would be an O(1) error
        whole5 = ub.inference.predict_volume(g, vol, patch=32, batch=5, shard=False)
        shard5 = ub.inference.predict_volume(g, vol, patch=32, batch=5)
        res["sharded_inference_batched_close"] = bool(((whole5 - shard5).norm() / whole5.norm()).item() < 1e-2)
        q.put((rank, res))
        dist.destroy_process_group()
    except Exception as ex:   # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, {"exception": traceback.format_exc()[-1500:]}))


def test_data_parallel_two_ranks_on_the_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank in (0, 1):
        assert "exception" not in res[rank], res[rank]["exception"]
        for k, v in res[rank].items():
            assert v is True, (rank, k, v)
