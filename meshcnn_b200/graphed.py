"""Whole training step as ONE CUDA graph.

A MeshCNN step at the reference's batch sizes is ~200 small kernels (36 MeshConv + norms + pools for the U-Net); on a
B200 each lasts a few microseconds, so a step driven launch by launch from Python is bound by the host.  Nothing in the
step needs the host: the live edge counts, the collapse order, the pooled topology, the MeshUnion groups and the Adam
step counter all stay in HBM.  ``GraphedTrainStep`` captures

    [topology blob + features + labels: pinned host -> device]  ->  forward (MeshConv / MeshPool / MeshUnpool ...)
    -> loss -> backward -> flat gradient all-reduce (NCCL, when world > 1) -> Adam

once and replays it.  Per step the host packs the next batch of meshes into the pinned staging buffers
(``MeshBatch.pack``) and calls ``run()``.

    step = GraphedTrainStep(net, criterion, reducer, optimizer, meshes0, x0, y0, E)
    for meshes, x, y in loader:
        loss = step(meshes, x, y)          # tensor on the device; .item() when the value is wanted
"""
import os

import numpy as np
import torch

from . import functional as Fn
from .mesh import MeshBatch

_CAPTURE_STREAM = {}     # device -> the stream every GraphedTrainStep warms up and captures on


class GraphedTrainStep:
    def __init__(self, net, criterion, reducer, optimizer, meshes, x, y, E, source='host', V_cap=0, ve_cap=0, warmup=3,
                 collective='auto'):
        """meshes / x / y: a first batch that fixes the shapes (x: [B, C, E] fp32 host array or tensor, y: labels).
        source='host': every replay starts with the host->device copies of the staged batch;
        source='device': the batch given here stays resident and every replay restores its pristine topology
        (benchmarking the device-side step alone)."""
        if source not in ('host', 'device'):
            raise ValueError(source)
        # Where the gradient all-reduce sits (world > 1).  'in-graph': one graph, the NCCL kernel is a node of it.
        # 'between': TWO graphs -- [copies, forward, backward] and [Adam, loss read-back] -- with the all-reduce launched
        # between them.  Measured on 2 x B200: the NCCL node costs the step 0.29-0.33 ms inside the graph of ~140 kernel nodes
        # but 47 us inside a graph of 8 (tools/ar_in_step_probe.py) and 39 us back to back -- i.e. ~2 us per node of the graph
        # it sits in, not NCCL time.  'auto' = 'between' when the reducer really reduces.
        if collective not in ('auto', 'in-graph', 'between'):
            raise ValueError(collective)
        multi = getattr(reducer, 'world', 1) > 1 and getattr(reducer, 'enabled', True)
        if collective == 'auto':
            collective = os.environ.get('MESHCNN_B200_COLLECTIVE', 'between' if multi else 'in-graph')
        self.collective = collective if multi else 'in-graph'
        self.net, self.criterion, self.reducer, self.opt, self.source = net, criterion, reducer, optimizer, source
        dev = reducer.flat.device
        self.device = dev
        x = torch.as_tensor(np.asarray(x)) if not torch.is_tensor(x) else x
        y = torch.as_tensor(np.asarray(y)) if not torch.is_tensor(y) else y
        self.x_host = torch.empty(x.shape, dtype=torch.float32).pin_memory()
        self.y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
        self.x_host.copy_(x)
        self.y_host.copy_(y)
        self.x_dev = torch.empty(x.shape, dtype=torch.float32, device=dev).requires_grad_(True)   # the reference feeds requires_grad inputs
        self.y_dev = torch.empty(y.shape, dtype=y.dtype, device=dev)
        meshes = list(meshes)
        if not V_cap:
            V_cap = int(1.25 * max(len(m._host_state()['vs']) for m in meshes)) + 8
        probe = MeshBatch(meshes, E, 'cpu', V_cap=V_cap, ve_cap=ve_cap)          # sizes only
        self.blob_host = torch.empty(probe.nbytes, dtype=torch.uint8).pin_memory()
        self.batch = MeshBatch(meshes, E, dev, V_cap=V_cap, ve_cap=ve_cap, staging=self.blob_host)
        with torch.no_grad():
            self.x_dev.copy_(self.x_host)
            self.y_dev.copy_(self.y_host)
        if source == 'device':
            self.batch.snapshot()
        self.h2d_bytes = self.blob_host.numel() + self.x_host.numel() * 4 + self.y_host.numel() * self.y_host.element_size()
        self.loss = None
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()   # the step's result, copied out by the graph itself
        self.done = torch.cuda.Event()
        self.graph = None
        self.graph_tail = None
        self._capture(warmup)

    # the work of one step, enqueued on the current stream
    def _enqueue(self):
        bt = self.batch
        Fn.begin_step(self.net)                     # weight images of every MeshConv on the side stream, under the H2D copies
        try:
            if self.source == 'host':
                bt.blob.copy_(self.blob_host[:bt.nbytes], non_blocking=True)
                with torch.no_grad():
                    self.x_dev.copy_(self.x_host, non_blocking=True)
                    self.y_dev.copy_(self.y_host, non_blocking=True)
                bt.rewind()
            else:
                bt.reset()
            self.reducer.zero()
            self.x_dev.grad = None
            out = self.net(self.x_dev, bt.meshes)
            loss = self.criterion(out, self.y_dev)
            loss.backward()
        finally:
            Fn.end_step()                           # also on an exception: a dangling step token would make later backward
                                                    # passes write straight into p.grad
        loss = loss.detach()                        # drop the autograd graph: it would keep the parameters' AccumulateGrad nodes
        if self.collective == 'in-graph':           # (and the stream they were created on) alive between objects
            self._enqueue_tail(loss)
        return loss

    def _enqueue_tail(self, loss):
        """What follows backward: gradient all-reduce (world > 1), optimiser step, loss read-back."""
        if self.collective == 'in-graph':
            self.reducer.allreduce()
        self.opt.step_state()
        self.loss_host.copy_(loss, non_blocking=True)

    def _capture(self, warmup):
        """Warm-up (allocator, lazy kernel configuration, NCCL) on a side stream, then capture.  The warm-up steps are real
        steps on the first batch; their effect on the training state -- parameters, Adam moments and step counter,
        BatchNorm running statistics -- is undone afterwards, so building a GraphedTrainStep does not train."""
        saved = self._training_state()
        cur = torch.cuda.current_stream()
        # warm-up AND capture on ONE side stream per device, shared by every object: autograd remembers the stream an
        # AccumulateGrad node was created on and the engine, at the end of backward, makes the caller's stream wait for each of
        # them -- even when the kernels wrote the gradient directly and handed autograd None.  With a different stream per
        # object the second capture waits on a stream that is not part of it ("dependency created on uncaptured work in another
        # stream", hit by the second GraphedTrainStep of a U-Net, which has no fc head to pull that stream into the capture).
        side = self._stream = _CAPTURE_STREAM.get(self.device)
        if side is None:
            side = self._stream = _CAPTURE_STREAM[self.device] = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        between = self.collective == 'between'
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                loss = self._enqueue()
                if between:
                    self.reducer.allreduce()
                    self._enqueue_tail(loss)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.batch.check()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = self._enqueue()
        self.graph_tail = None
        if between:
            self.graph_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_tail, pool=self.graph.pool(), stream=side):
                self._enqueue_tail(self.loss)
        self.warmup_steps = max(1, warmup)
        self._restore_training_state(saved)
        torch.cuda.synchronize()

    def _training_state(self):
        opt = self.opt
        st = {'buffers': [b.detach().clone() for b in self.net.buffers()], 'step_count': getattr(opt, 'step_count', None)}
        for k in ('flat_p', 'exp_avg', 'exp_avg_sq', 'state'):
            t = getattr(opt, k, None)
            st[k] = None if t is None else t.detach().clone()
        if st['flat_p'] is None:                     # not a FlatAdam: keep the parameters themselves
            st['params'] = [p.detach().clone() for p in self.net.parameters()]
        return st

    def _restore_training_state(self, st):
        with torch.no_grad():
            for b, v in zip(self.net.buffers(), st['buffers']):
                b.copy_(v)
            for k in ('flat_p', 'exp_avg', 'exp_avg_sq', 'state'):
                if st[k] is not None:
                    getattr(self.opt, k).copy_(st[k])
            if st.get('params') is not None:
                for p, v in zip(self.net.parameters(), st['params']):
                    p.copy_(v)
        if st['step_count'] is not None:
            self.opt.step_count = st['step_count']

    def stage(self, meshes, x, y):
        """Host side of a step: pack the next batch into the pinned staging buffers (no device work).  Waits for the previous
        replay of THIS graph first: its host->device copy nodes read these buffers (PipelinedTrainStep alternates two graphs
        so that this wait is normally over before it starts), and looks at the collapse status that replay left behind."""
        self.done.synchronize()
        self.batch.check_captured_status()
        self.batch.pack(meshes)
        self.x_host.copy_(torch.as_tensor(x) if not torch.is_tensor(x) else x)
        self.y_host.copy_(torch.as_tensor(y) if not torch.is_tensor(y) else y)

    def run(self):
        """Replay the captured step; returns the (device) loss tensor of the graph's static output."""
        self.opt.sync_lr()
        self.graph.replay()
        if self.graph_tail is not None:
            self.reducer.allreduce()                # one NCCL launch between the two graphs
            self.graph_tail.replay()
        self.done.record()
        self.batch.version += 1                     # host mirrors of the meshes refresh lazily from the new state
        return self.loss

    def close(self):
        """Release the captured graphs (before the NCCL communicator they may reference is destroyed)."""
        self.graph = None
        self.graph_tail = None

    def result(self):
        """Loss of the most recent run() as a Python float; waits for THAT replay only (later work keeps running)."""
        self.done.synchronize()
        self.batch.check_captured_status()
        return float(self.loss_host)

    def __call__(self, meshes, x, y):
        if self.source != 'host':
            raise RuntimeError("a source='device' step replays its resident batch: call run()")
        self.stage(meshes, x, y)
        return self.run()


class PipelinedTrainStep:
    """Two GraphedTrainSteps over the same network / optimiser with their own staging buffers: while the GPU replays step k
    from slot A, the host packs the meshes of step k+1 into slot B.  ``submit`` returns the slot; ``slot.result()`` is the
    loss of that step (one event wait, no device-wide synchronisation)."""

    def __init__(self, net, criterion, reducer, optimizer, meshes, x, y, E, **kw):
        self.slots = [GraphedTrainStep(net, criterion, reducer, optimizer, meshes, x, y, E, source='host', **kw)]
        self.slots.append(GraphedTrainStep(net, criterion, reducer, optimizer, self._detached(meshes), x, y, E, source='host', **kw))
        self.k = 0
        self.h2d_bytes = self.slots[0].h2d_bytes

    @staticmethod
    def _detached(meshes):
        for m in meshes:
            m._batch = None
        return meshes

    def submit(self, meshes, x, y):
        slot = self.slots[self.k & 1]
        self.k += 1
        slot.stage(meshes, x, y)                    # waits for the replay that last read this slot's staging buffers
        slot.run()
        return slot
