"""Oracle networks vs end-to-end outputs of the reference's own models/networks.py (tests/golden/net_*.npz)."""
import functools

import numpy as np
import torch

from oracle import meshcnn_oracle as O
from oracle import networks_oracle as ON
from tests.helpers import close, golden, prepared, rel_err


def _load(net, d):
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith('sd/')}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def _meshes(d, hold):
    return [O.OracleMesh(prepared(d['M%d_vs0' % i], d['M%d_faces' % i]), hold_history=hold) for i in range(3)]


def _check(net, d, loss_fn, hold):
    _load(net, d)
    x = torch.from_numpy(d['x']).requires_grad_(True)
    out = net(x, _meshes(d, hold))
    loss = loss_fn(out, torch.from_numpy(d['labels']))
    loss.backward()
    assert rel_err(out.detach().numpy(), d['out']) < 1e-4
    assert abs(float(loss.detach()) - float(d['loss'])) < 1e-5
    assert rel_err(x.grad.numpy(), d['dx']) < 1e-3
    for k, p in net.named_parameters():
        assert close(p.grad.numpy(), d['grad/' + k], 1e-3, 2e-6), k


def test_mconvnet_matches_reference():
    d = golden('net_mconvnet_tiny')
    norm = functools.partial(torch.nn.GroupNorm, 4)
    for dense in (False, True):
        net = ON.OMeshConvNet(norm, 5, [8, 16], 4, 270, [200, 150], 10, 1, dense=dense)
        _check(net, d, torch.nn.CrossEntropyLoss(), False)


def test_meshunet_matches_reference():
    d = golden('net_meshunet_tiny')
    net = ON.OMeshEncoderDecoder([270, 200, 150], [5, 8, 16, 32], [32, 16, 8, 3], blocks=1)
    _check(net, d, torch.nn.CrossEntropyLoss(ignore_index=-1), True)


def test_state_dict_keys_match_product_networks():
    """meshcnn_b200.networks must expose exactly the reference's parameter names; construction only."""
    from meshcnn_b200 import networks as N
    d = golden('net_mconvnet_tiny')
    norm = N.get_norm_layer('group', 4)
    net = N.MeshConvNet(norm, 5, [8, 16], 4, 270, [200, 150], 10, 1)
    want = {k[3:]: d[k].shape for k in d.files if k.startswith('sd/')}
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    d = golden('net_meshunet_tiny')
    net = N.MeshEncoderDecoder([270, 200, 150], [5, 8, 16, 32], [32, 16, 8, 3], blocks=1)
    want = {k[3:]: tuple(d[k].shape) for k in d.files if k.startswith('sd/')}
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == want
