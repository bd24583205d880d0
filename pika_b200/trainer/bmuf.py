"""BMUF (block-wise model update filtering) -- drop-in for trainer/bmuf.py:BmufTrainer (reference).

Same constructor and methods (``update_and_sync() -> SUCCESS|STOP``, ``sum_reduce``, ``broadcast``).
B200 mapping: the reference's ``reduce(delta -> rank 0)`` / rank-0 update / ``broadcast(param)``
(trainer/bmuf.py:83-98) is one NCCL all-reduce over NVLink 5 / NVSwitch of the flat delta followed by
the identical fused block-momentum update on every rank (``delta_prev`` is replicated instead of living
on rank 0 only); parameters are views of the flat buffer, so there are no flatten / un-flatten copies.
The NaN guard is made collective (the reference lets only the ranks that see NaN return early and the
others hang in ``broadcast`` -- SURVEY.md section 5).
"""
import torch
import torch.distributed as dist

from .. import engine
from .. import kernels as K
from .flat import FlatParams

SUCCESS = 1
STOP = 0


class BmufTrainer():
    """
    Args (as in the reference):
        master_node (int), rank (int), world_size (int), model (nn.Module),
        block_momentum (float), block_lr (float)
    Extra keyword: ``flat`` -- an existing FlatParams of ``model`` (created when omitted);
    ``backend`` -- "nccl" (default, as hard-coded in the reference) or "gloo" for CPU-side tests.
    """

    def __init__(self, master_node, rank, world_size, model, block_momentum, block_lr, flat=None, backend="nccl", ops=None):
        # ``ops``: object with bmuf_delta / bmuf_update / absmax; defaults to the CUDA kernels.  The CPU-side
        # gloo test injects the numpy oracle here to exercise the collective protocol without a GPU.
        self.ops = ops if ops is not None else K
        self.master_node, self.rank, self.world_size = master_node, rank, world_size
        self.model, self.block_momentum, self.block_lr = model, block_momentum, block_lr
        if world_size > 1 and not dist.is_initialized():
            dist.init_process_group(backend=backend, init_method="env://")
        self.flat = flat if flat is not None else FlatParams(model)
        self.param = self.flat.data.clone()              # the global (block) model
        if world_size > 1:
            dist.broadcast(tensor=self.param, src=master_node)
            self.flat.data.copy_(self.param)
            engine.invalidate_weights()
        self.delta_prev = torch.zeros_like(self.param)
        self.delta = torch.zeros_like(self.param)
        if world_size > 1:
            # the first all-reduce of a communicator sets up its channels / buffers for this message size (measured on 8 B200s: ~150 ms
            # against 2.3 ms for every later block sync): pay that at construction, next to the parameter broadcast, not in the first block
            dist.all_reduce(self.delta, op=dist.ReduceOp.SUM)
        self.health = torch.zeros(2, dtype=torch.float32, device=self.param.device)   # [absmax, unused]
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.param.device)

    def update_and_sync(self):
        """one block sync: returns SUCCESS if numerics are healthy on every rank, STOP otherwise"""
        self.ops.bmuf_delta(self.param, self.flat.data, self.delta)
        if self.world_size > 1:
            dist.all_reduce(self.delta, op=dist.ReduceOp.SUM)
        self.nan_flag.zero_()
        self.ops.absmax(self.delta, self.health[:1], self.nan_flag)     # NaNs propagate through the sum: every rank sees them
        if int(self.nan_flag.item()) != 0:
            return STOP
        self.ops.bmuf_update(self.param, self.flat.data, self.delta_prev, self.delta, self.world_size, self.block_momentum, self.block_lr)
        engine.invalidate_weights()
        return SUCCESS

    def broadcast(self, tensor):
        """broadcast interface for trainer"""
        if self.world_size > 1:
            dist.broadcast(tensor=tensor, src=self.master_node)

    def sum_reduce(self, tensor):
        """sumreduce interface for trainer (result valid on the master node, as in the reference)"""
        if self.world_size > 1:
            dist.reduce(tensor=tensor, dst=self.master_node)
