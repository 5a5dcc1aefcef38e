"""Fused classifier head (AvgPool1d -> fc1 -> relu -> fc2; models/networks.py:152-156 of the reference) vs the same ops in plain
PyTorch fp32: logits and every gradient within 1e-5 relative."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.helpers import close

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('B,E,C,H,K', [(64, 210, 256, 100, 22), (16, 180, 256, 100, 30), (3, 37, 64, 17, 5), (5, 50, 1028, 300, 40),
                                         (2, 9, 8, 1, 1)])
@pytest.mark.parametrize('bias', [True, False])
def test_head_matches_torch(B, E, C, H, K, bias):
    from meshcnn_b200.functional import ClassifierHeadFn
    torch.manual_seed(B * 1000 + C)
    fc1, fc2 = nn.Linear(C, H, bias=bias).to(DEV), nn.Linear(H, K, bias=bias).to(DEV)
    x = torch.randn(B, E, C, device=DEV)
    x[:, E - 3:, :] = 0.0                                    # zero-padded slots are part of the average, as in the reference
    y = torch.randint(0, K, (B,), device=DEV)
    x1 = x.clone().requires_grad_(True)
    out1 = ClassifierHeadFn.apply(x1, fc1.weight, fc1.bias, fc2.weight, fc2.bias)
    F.cross_entropy(out1, y).backward()
    got = [out1.detach(), x1.grad] + [p.grad.clone() for p in list(fc1.parameters()) + list(fc2.parameters())]
    for p in list(fc1.parameters()) + list(fc2.parameters()):
        p.grad = None
    x2 = x.clone().requires_grad_(True)
    # the reference's ops on its own layout: x [B, C, E] -> AvgPool1d(E) -> view -> relu(fc1) -> fc2
    pooled = nn.AvgPool1d(E)(x2.permute(0, 2, 1)).view(-1, C)
    out2 = fc2(F.relu(fc1(pooled)))
    F.cross_entropy(out2, y).backward()
    ref = [out2.detach(), x2.grad] + [p.grad.clone() for p in list(fc1.parameters()) + list(fc2.parameters())]
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        assert close(g.cpu().numpy(), r.cpu().numpy(), 1e-5)


def test_head_writes_gradients_in_place_inside_a_step():
    """Inside a step opened with begin_step() the gradients land in p.grad directly and autograd gets None (no AccumulateGrad
    add kernels); the values are those of the eager path."""
    from meshcnn_b200 import functional as Fn
    from meshcnn_b200.functional import ClassifierHeadFn
    torch.manual_seed(3)
    B, E, C, H, K = 8, 60, 64, 20, 7
    fc1, fc2 = nn.Linear(C, H).to(DEV), nn.Linear(H, K).to(DEV)
    params = list(fc1.parameters()) + list(fc2.parameters())
    x = torch.randn(B, E, C, device=DEV)
    y = torch.randint(0, K, (B,), device=DEV)
    F.cross_entropy(ClassifierHeadFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias), y).backward()
    ref = [p.grad.clone() for p in params]
    for p in params:
        p.grad = torch.zeros_like(p)
    bufs = [p.grad for p in params]
    Fn.STEP_TOKEN = 12345
    try:
        F.cross_entropy(ClassifierHeadFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias), y).backward()
    finally:
        Fn.STEP_TOKEN = -1
    for p, b, r in zip(params, bufs, ref):
        assert p.grad is b
        assert close(b.cpu().numpy(), r.cpu().numpy(), 1e-6)
