"""The drop-in registration: the reference's unmodified models/networks.py must bind to the CUDA-backed layer classes
and build both networks with the reference's state-dict keys.  Construction only (no GPU here); skipped where the
reference tree is absent (the GPU box)."""
import os
import subprocess
import sys

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, types, numpy as np
sys.path.insert(0, %r)
from meshcnn_b200 import dropin
dropin.install(%r)
import models.networks as nets                       # the reference's own file
import meshcnn_b200.mesh_conv as mc, meshcnn_b200.mesh_pool as mp, meshcnn_b200.mesh_unpool as mu
assert nets.MeshConv is mc.MeshConv and nets.MeshPool is mp.MeshPool and nets.MeshUnpool is mu.MeshUnpool
from data.classification_data import Mesh as DataMesh
import meshcnn_b200.mesh as mm
assert DataMesh is mm.Mesh
opt = types.SimpleNamespace(norm='group', num_groups=4, pool_res=[200, 150], fc_n=10, resblocks=1)
net = nets.define_classifier(5, [8, 16], 270, 4, opt, [], 'mconvnet', 'normal', 0.02)
keys = sorted(net.state_dict())
assert 'conv0.conv0.conv.weight' in keys and 'conv1.bn1.weight' in keys and 'fc2.bias' in keys, keys
assert tuple(net.state_dict()['conv0.conv0.conv.weight'].shape) == (8, 5, 1, 5)
opt = types.SimpleNamespace(norm='group', num_groups=4, pool_res=[200, 150], fc_n=10, resblocks=1)
unet = nets.define_classifier(5, [8, 16, 32], 270, 3, opt, [], 'meshunet', 'normal', 0.02)
assert 'decoder.final_conv.conv1.conv.bias' in unet.state_dict()
# and our own networks module exposes exactly the same keys
from meshcnn_b200 import networks as N
mine = N.define_classifier(5, [8, 16, 32], 270, 3, opt, [], 'meshunet', 'normal', 0.02)
assert sorted(mine.state_dict()) == sorted(unet.state_dict())
print('DROPIN-OK')
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_reference_networks_bind_to_cuda_layers():
    r = subprocess.run([sys.executable, '-c', CODE % (ROOT, REF)], capture_output=True, text=True, timeout=300)
    assert 'DROPIN-OK' in r.stdout, r.stdout + r.stderr
