"""Flat-buffer Adam: the parameters of a network become views into ONE fp32 buffer (as their gradients already are in
ddp.FlatGradReducer), so the optimiser step of mesh_classifier.py:38 (torch.optim.Adam) is a single kernel
(mcnn_adam_flat) instead of a multi-tensor launch per group."""
import torch

from . import _lib
from . import profiler


class FlatAdam:
    def __init__(self, reducer, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.reducer = reducer
        self.params = reducer.params
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        n = reducer.flat.numel()
        dev = reducer.flat.device
        if dev.type != 'cuda':
            raise RuntimeError('FlatAdam runs on CUDA only')
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                o += k
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_count = 0
        self.param_groups = [{'lr': lr}]                       # so lr schedulers / logging can read and set it
        # device-resident copy of the hyper-parameters + step counter: the launch sequence is then CUDA-graph capturable
        self.state = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self._state_lr = lr

    def zero_grad(self, set_to_none=False):
        self.reducer.zero()

    def sync_lr(self):
        """Push a learning rate changed through param_groups into the device state (call outside graph capture)."""
        lr = float(self.param_groups[0]['lr'])
        if lr != self._state_lr:
            self.state[0:1].fill_(lr)
            self._state_lr = lr

    def step_state(self):
        """The same update driven entirely by device state (lr, step count): safe to capture and replay."""
        self.step_count += 1
        with profiler.span('adam_flat', n=self.flat_p.numel()):
            _lib.check(_lib.lib().mcnn_adam_flat_state(self.flat_p.data_ptr(), self.reducer.flat.data_ptr(), self.exp_avg.data_ptr(),
                                                       self.exp_avg_sq.data_ptr(), self.flat_p.numel(), self.state.data_ptr(),
                                                       _lib.current_stream()), 'mcnn_adam_flat_state')

    def step(self):
        self.step_count += 1
        self.state[5:6].add_(1.0)                              # keep the device counter in step with the host one
        lr = self.param_groups[0]['lr']
        with profiler.span('adam_flat', n=self.flat_p.numel()):
            _lib.check(_lib.lib().mcnn_adam_flat(self.flat_p.data_ptr(), self.reducer.flat.data_ptr(), self.exp_avg.data_ptr(),
                                                 self.exp_avg_sq.data_ptr(), self.flat_p.numel(), lr, self.betas[0], self.betas[1],
                                                 self.eps, self.weight_decay, self.step_count, _lib.current_stream()),
                       'mcnn_adam_flat')
