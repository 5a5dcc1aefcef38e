"""Fused norm blocks and the flat Adam kernel vs plain PyTorch fp32 references of the same ops (1e-4 relative)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.helpers import close

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ref(norm, x, a, r, pre_relu, post_relu):
    """torch reference on the logical [B, C, E, 1] tensors."""
    def lg(t):
        return t.permute(0, 2, 1).unsqueeze(-1)
    z = lg(x) if a is None else lg(x) + lg(a)
    if pre_relu:
        z = F.relu(z)
    y = norm(z)
    if r is not None:
        y = y + lg(r)
    if post_relu:
        y = F.relu(y)
    return y.squeeze(-1).permute(0, 2, 1)


@pytest.mark.parametrize('kind,C', [('batch', 64), ('group', 128), ('instance', 32), ('group', 256), ('instance', 8), ('batch', 256)])
@pytest.mark.parametrize('variant', ['plain', 'pre', 'post'])
@pytest.mark.parametrize('B,E', [(6, 750), (16, 300), (9, 1400)], ids=['three_kernels', 'one_launch', 'one_launch_narrow'])
def test_fused_norm_matches_torch(kind, C, variant, B, E):
    """B = 6: statistics / finalize / apply kernels; B >= 8 with per-mesh statistics (group, instance): the one-launch form
    that keeps the mesh's slab in shared memory (1400 rows: 16-channel chunks)."""
    from meshcnn_b200.functional import fused_norm
    torch.manual_seed(C + len(variant))
    mk = {'batch': lambda: nn.BatchNorm2d(C), 'group': lambda: nn.GroupNorm(16 if C >= 16 else 4, C),
          'instance': lambda: nn.InstanceNorm2d(C)}[kind]
    n1, n2 = mk().to(DEV).train(), mk().to(DEV).train()
    if kind != 'instance':
        with torch.no_grad():
            n1.weight.uniform_(0.5, 1.5); n1.bias.uniform_(-0.3, 0.3)
        n2.load_state_dict(n1.state_dict())
    x = (torch.randn(B, E, C, device=DEV) * 1.7 + 0.4)
    aux = torch.randn(B, E, C, device=DEV)
    a = aux if variant == 'pre' else None
    r = aux if variant == 'post' else None
    pre_relu = variant == 'pre'
    post_relu = variant != 'plain'
    xs = [x.clone().requires_grad_(True) for _ in range(2)]
    auxs = [aux.clone().requires_grad_(True) for _ in range(2)]
    y1 = fused_norm(xs[0], n1, pre_add=auxs[0] if a is not None else None, pre_relu=pre_relu,
                    post_add=auxs[0] if r is not None else None, post_relu=post_relu)
    assert y1 is not None
    y2 = _ref(n2, xs[1], auxs[1] if a is not None else None, auxs[1] if r is not None else None, pre_relu, post_relu)
    assert close(y1.detach().cpu().numpy(), y2.detach().cpu().numpy(), 1e-4)
    dy = torch.randn_like(y1)
    y1.backward(dy)
    y2.backward(dy)
    assert close(xs[0].grad.cpu().numpy(), xs[1].grad.cpu().numpy(), 1e-4, 1e-6)
    if variant != 'plain':
        assert close(auxs[0].grad.cpu().numpy(), auxs[1].grad.cpu().numpy(), 1e-4, 1e-6)
    if kind != 'instance':
        assert close(n1.weight.grad.cpu().numpy(), n2.weight.grad.cpu().numpy(), 1e-4, 1e-5)
        assert close(n1.bias.grad.cpu().numpy(), n2.bias.grad.cpu().numpy(), 1e-4, 1e-5)
    if kind == 'batch':
        assert close(n1.running_mean.cpu().numpy(), n2.running_mean.cpu().numpy(), 1e-5, 1e-7)
        assert close(n1.running_var.cpu().numpy(), n2.running_var.cpu().numpy(), 1e-5, 1e-7)
        assert int(n1.num_batches_tracked) == int(n2.num_batches_tracked) == 1
        # eval mode uses the running statistics
        n1.eval(); n2.eval()
        with torch.no_grad():
            e1 = fused_norm(x, n1, post_relu=True)
            e2 = _ref(n2, x, None, None, False, True)
        assert close(e1.cpu().numpy(), e2.cpu().numpy(), 1e-4)


def test_flat_adam_matches_torch_adam():
    from meshcnn_b200 import ddp
    from meshcnn_b200.optim import FlatAdam
    torch.manual_seed(3)
    net1 = nn.Sequential(nn.Linear(37, 50), nn.ReLU(), nn.Linear(50, 11)).to(DEV)
    net2 = nn.Sequential(nn.Linear(37, 50), nn.ReLU(), nn.Linear(50, 11)).to(DEV)
    net2.load_state_dict(net1.state_dict())
    red = ddp.FlatGradReducer(net1)
    opt1 = FlatAdam(red, lr=1e-3, betas=(0.9, 0.999))
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-3, betas=(0.9, 0.999))
    for it in range(5):
        x = torch.randn(16, 37, device=DEV)
        y = torch.randint(0, 11, (16,), device=DEV)
        opt1.zero_grad()
        F.cross_entropy(net1(x), y).backward()
        assert red.check_views()
        opt1.step()
        opt2.zero_grad()
        F.cross_entropy(net2(x), y).backward()
        opt2.step()
    for p1, p2 in zip(net1.parameters(), net2.parameters()):
        assert close(p1.detach().cpu().numpy(), p2.detach().cpu().numpy(), 1e-5, 1e-7)


@pytest.mark.parametrize('kind,C,E', [('group', 64, 750), ('group', 256, 300), ('instance', 128, 600), ('group', 32, 1400)])
def test_one_launch_norm_matches_three_kernel_form(kind, C, E):
    """Same statistics, same formulas: the two forms must agree to fp32 rounding of the partial sums."""
    from meshcnn_b200 import _lib
    from meshcnn_b200.functional import fused_norm
    torch.manual_seed(C + E)
    B = 12
    norm = (nn.GroupNorm(16, C) if kind == 'group' else nn.InstanceNorm2d(C)).to(DEV).train()
    if kind == 'group':
        with torch.no_grad():
            norm.weight.uniform_(0.5, 1.5); norm.bias.uniform_(-0.3, 0.3)
    x = torch.randn(B, E, C, device=DEV) * 1.3 - 0.2
    aux = torch.randn(B, E, C, device=DEV)
    dy = torch.randn(B, E, C, device=DEV)
    L = _lib.lib()
    res = []
    try:
        for on in (0, 1):
            L.mcnn_debug_set_norm_fused(on)
            xs, ax = x.clone().requires_grad_(True), aux.clone().requires_grad_(True)
            norm.zero_grad()
            y = fused_norm(xs, norm, pre_add=ax, pre_relu=True, post_relu=True)
            y.backward(dy)
            torch.cuda.synchronize()
            res.append([y.detach().clone(), xs.grad.clone(), ax.grad.clone()] + ([norm.weight.grad.clone(), norm.bias.grad.clone()] if kind == 'group' else []))
    finally:
        L.mcnn_debug_set_norm_fused(1)
    for p3, p1 in zip(*res):
        assert close(p1.cpu().numpy(), p3.cpu().numpy(), 2e-6, 1e-7)
