"""BMUF protocol across 2 ranks on CPU (gloo): the drop-in BmufTrainer's collective sequence
(initial broadcast, all-reduce of the block delta, replicated block-momentum update, collective NaN stop,
sum_reduce/broadcast helpers) against the oracle of trainer/bmuf.py:76-100 AND against the parameter trajectory of the
reference's own BmufTrainer run on 2 gloo ranks in the build container (tests/golden/bmuf_2rank.npz, make_golden.py:golden_bmuf).  The element-wise kernels are
CUDA-only, so this host-logic test injects numpy implementations of the three ops (oracle/train.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyOps:
    @staticmethod
    def bmuf_delta(glob, local, delta):
        delta.copy_(glob - local)

    @staticmethod
    def absmax(x, out, nan_flag):
        out[0] = x.abs().max() if not torch.isnan(x).any() else float("nan")
        if torch.isnan(x).any():
            nan_flag[0] = 1

    @staticmethod
    def bmuf_update(glob, local, delta_prev, delta_sum, world, bm, blr):
        d = delta_sum / float(world)
        delta_prev.copy_(bm * delta_prev + blr * (1 - bm) * d)
        glob.sub_((1 + bm) * delta_prev)
        local.copy_(glob)


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from oracle import train as ot
    from pika_b200.trainer.bmuf import BmufTrainer, SUCCESS, STOP
    torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights: rank 0's must win
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    w0 = [torch.nn.utils.parameters_to_vector(model.parameters()).detach().clone()]
    bm, blr = 0.9, 1.0
    tr = BmufTrainer(0, rank, world, model, bm, blr, backend="gloo", ops=NumpyOps)
    flat0 = tr.param.clone()
    # every rank now holds rank 0's initial parameters, and the model's parameters are views of the flat buffer
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, tr.flat.data.clone())
    same_init = all(torch.equal(g, gathered[0]) for g in gathered)
    views_ok = all(p.data_ptr() >= tr.flat.data.data_ptr() for p in model.parameters())
    # two block syncs with rank-specific local updates, checked against the oracle formula
    glob = flat0.numpy().astype(np.float64).copy()
    dprev = np.zeros_like(glob)
    ok = True
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bmuf_2rank.npz"))["params"]      # [1 + syncs, n] reference trajectory
    pvec = lambda: torch.nn.utils.parameters_to_vector(model.parameters()).detach().numpy()   # noqa: E731  (flat buffer minus alignment padding)
    ok &= np.array_equal(pvec(), gold[0])                                                 # rank 0's initial weights everywhere
    for it in range(3):
        g = torch.Generator().manual_seed(7 * it + rank)
        with torch.no_grad():
            noise = 0.01 * torch.randn(gold.shape[1], generator=g)          # same draw as the reference run: one value per parameter
            off = 0
            for prm in model.parameters():                                  # in place: parameters are views of the flat buffer
                prm.add_(noise[off:off + prm.numel()].view_as(prm))
                off += prm.numel()
        locals_ = [torch.zeros_like(flat0) for _ in range(world)]
        dist.all_gather(locals_, tr.flat.data.clone())
        assert tr.update_and_sync() == SUCCESS
        glob, dprev = ot.bmuf_update(glob, [l.numpy().astype(np.float64) for l in locals_], dprev, bm, blr)
        ok &= np.allclose(tr.param.numpy(), glob, atol=1e-6) and torch.equal(tr.param, tr.flat.data)
        # reference: reduce to rank 0 -> update -> broadcast; here: all-reduce -> replicated update.  Same fp32 formula; the only
        # difference is gloo's summation order over 2 ranks (commutative), so the trajectories agree to the last few ulps
        ok &= np.allclose(pvec(), gold[it + 1], rtol=0, atol=2e-7)
    # helper collectives
    t = torch.tensor([float(rank + 1), 10.0])
    tr.sum_reduce(t)
    tr.broadcast(t)
    helpers_ok = abs(t[0].item() - sum(range(1, world + 1))) < 1e-6
    # collective NaN stop: only rank 1 diverges, EVERY rank must return STOP (no hang)
    if rank == 1:
        with torch.no_grad():
            tr.flat.data[3] = float("nan")
    stop = tr.update_and_sync()
    q.put((rank, same_init, views_ok, bool(ok), helpers_ok, stop == STOP))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bmuf_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
    for r in res:
        assert all(r[1:]), r


def worker_empty(rank, world, port, q):
    """rank 1 receives an EMPTY loader item exactly at a sync index; it must still enter the block sync (reference loop
    trainer/train_transducer_bmuf_otfaug.py:112-123 syncs on every `num_done % sync_period == 0`, data or not)"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import types
    from pika_b200.trainer.bmuf import BmufTrainer
    from pika_b200.trainer.step import TrainStep
    torch.manual_seed(5)
    model = torch.nn.Linear(6, 4)
    tr = BmufTrainer(0, rank, world, model, 0.9, 1.0, backend="gloo", ops=NumpyOps)
    resets = []
    opt = types.SimpleNamespace(reset=lambda lr: resets.append(lr))
    args = types.SimpleNamespace(sync_period=2, initial_lr=1e-3, final_lr=1e-4, epoch=0, num_batches_per_epoch=10, num_epochs=2)
    step = TrainStep(model, args, None, tr, opt)
    for item in range(5):                              # sync indices: 2 and 4
        empty = (rank == 1 and item in (2, 3))         # rank 1: empty at a sync index and at a plain index
        if not empty:
            with torch.no_grad():
                model.weight.add_(0.01 * (rank + 1) * (item + 1))     # stands in for forward/backward/SGD
            step.end_of_item()
        else:
            step.skip()
    mine = tr.flat.data.clone()
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    q.put((rank, step.num_done == 5, len(resets) == 2, torch.equal(gathered[0], gathered[1]), torch.equal(tr.param, tr.flat.data)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_empty_batch_on_one_rank_still_syncs():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker_empty, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
    for r in res:
        assert all(r[1:]), r
