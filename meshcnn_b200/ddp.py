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
    def __init__(self, net, local_items=None, global_items=None, early_tail=False):
        self.params = [p for p in net.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        self.offsets = {}
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)          # gradient-as-bucket-view
            self.offsets[id(p)] = o
            o += p.numel()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # uneven shards (e.g. 12 meshes on 8 GPUs): weight each rank's mean gradient by its share of the batch
        self.scale = 1.0 / self.world if local_items is None else float(local_items) / float(global_items)
        self.equal_shards = local_items is None
        self.enabled = True       # False: allreduce() is a no-op (scaling diagnostics: the step's device time without the collective)
        # Early tail bucket.  The late layers hold most of the parameters (cubes: conv2 + conv3 + fc = 89 %) and their gradients
        # are complete long before backward ends, so their all-reduce can run on a side stream under the rest of backward;
        # only the small head bucket is reduced after it.  Used inside a step opened with functional.begin_step() (the
        # kernels write gradients straight into this buffer there, so "launched" means "final"); the trigger is the MeshConv
        # weight at which the tail starts, fired from MeshConvFn.backward once that layer's gradient kernels are enqueued.
        # OFF by default: measured on 2 x B200 it does not shorten the step (3.05 -> 3.07 ms) -- what the all-reduce costs a
        # captured step is not NCCL time (see GraphedTrainStep: the collective now sits BETWEEN two graphs).
        self.split, self.trigger, self._tail_pending, self._comm = None, None, False, None
        self._conv_weights, self._seen = [], set()
        if early_tail and self.world > 1 and dev.type == 'cuda':
            self._plan_tail(net)
            from . import functional as Fn
            Fn.GRAD_REDUCER = self

    def _plan_tail(self, net):
        from .mesh_conv import MeshConv
        convs = [m.conv.weight for m in net.modules() if isinstance(m, MeshConv) and m.conv.weight.requires_grad]
        self._conv_weights = convs
        n = self.flat.numel()
        best = None
        for w in convs:                                  # the first layer (in parameter order) behind the first 5 % of the buffer:
            off = self.offsets[id(w)]                    # the head bucket stays small, and the early layers' backward (large edge
            if off >= 0.05 * n and (best is None or off < self.offsets[id(best)]):   # counts) is left to hide the tail's all-reduce
                best = w
        if best is not None:
            self.trigger, self.split = best, self.offsets[id(best)]

    def zero(self):
        self.flat.zero_()
        self._seen.clear()
        self._tail_pending = False

    def _reduce(self, t):
        if self.equal_shards and t.is_cuda and dist.get_backend() == 'nccl':
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            t.mul_(self.scale)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def conv_backward_launched(self, weight):
        """Called by MeshConvFn.backward (inside a begin_step() step) once the gradient kernels of a layer are enqueued."""
        if self.split is None or self.world == 1 or not self.enabled:
            return
        self._seen.add(id(weight))
        if weight is not self.trigger or self._tail_pending:
            return
        # every MeshConv of the tail must have been through backward already (true for a feed-forward net and for the U-Net,
        # whose decoder precedes the encoder in backward); otherwise fall back to one all-reduce at the end
        if not all(id(w) in self._seen for w in self._conv_weights if self.offsets[id(w)] >= self.split):
            return
        cur = torch.cuda.current_stream()
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.flat.device)
        self._comm.wait_stream(cur)
        with torch.cuda.stream(self._comm):
            self._reduce(self.flat[self.split:])
        self._tail_pending = True

    def allreduce(self):
        if self.world == 1 or not self.enabled or os.environ.get('MESHCNN_B200_SKIP_ALLREDUCE') == '1':     # the latter two: scaling diagnostics only
            return
        if self._tail_pending:
            self._reduce(self.flat[:self.split])
            torch.cuda.current_stream().wait_stream(self._comm)
            self._tail_pending = False
        else:
            self._reduce(self.flat)

    def check_views(self):
        """True while autograd still accumulates into the flat buffer (it does unless someone sets grads to None)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)
