"""Data parallelism for the mesh batch: one process per GPU, meshes sharded contiguously, ONE flat gradient
allreduce per step.  The reference's nn.DataParallel cannot shard the mesh list; this
is new code, not a port.

The gradients of all parameters are views into one flat fp32 buffer, so zeroing them is one memset, the allreduce is
one NCCL call over NVLink / NVSwitch on a 5-9 MB bucket (latency-bound: one bucket, not many), and no gather / scatter
copies are needed.  Works with the gloo backend on CPU tensors as well (tests/test_ddp_cpu.py)."""
import os

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of a batch of n_items for `rank`; the first n_items % world ranks get one more."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_parameters(net, src=0):
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, src)


def enable_sync_batchnorm(on=True, local_items=None, global_items=None):
    """Batch statistics of every BatchNorm2d (MResConv.bn, networks.py:167) over the rows of ALL ranks instead of the local
    shard: the forward all-reduces (sum, sum of squares, count) per channel, the backward (sum dy, sum dy * xhat).
    For uneven shards pass this rank's and the global mesh count, as to FlatGradReducer."""
    from . import functional as Fn
    Fn.SYNC_BN = bool(on)
    Fn.SYNC_BN_WEIGHT = None if local_items is None else float(local_items) / float(global_items)


class FlatGradReducer:
    def __init__(self, net, local_items=None, global_items=None):
        self.params = [p for p in net.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)          # gradient-as-bucket-view
            o += p.numel()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # uneven shards (e.g. 12 meshes on 8 GPUs): weight each rank's mean gradient by its share of the batch
        self.scale = 1.0 / self.world if local_items is None else float(local_items) / float(global_items)
        self.enabled = True       # False: allreduce() is a no-op (scaling diagnostics: the step's device time without the collective)

    def zero(self):
        self.flat.zero_()

    def allreduce(self):
        if self.world == 1 or not self.enabled or os.environ.get('MESHCNN_B200_SKIP_ALLREDUCE') == '1':     # the latter two: scaling diagnostics only
            return
        if self.scale != 1.0:
            self.flat.mul_(self.scale)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def check_views(self):
        """True while autograd still accumulates into the flat buffer (it does unless someone sets grads to None)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)
