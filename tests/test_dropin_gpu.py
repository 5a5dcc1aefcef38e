"""The drop-in claim, exercised end to end on the GPU: the reference's OWN, unmodified train.py / test.py / models/networks.py
/ models/mesh_classifier.py / data loaders run on the CUDA layers through meshcnn_b200.dropin (the reference tree is
/root/reference in the build container, the copy build() stages into oracle/_ref/reference on the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import reference_arm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(reference_arm.find_reference() is None, reason='reference tree not present')]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NETS = r'''
import sys, types, numpy as np, torch
sys.path.insert(0, %(root)r)
from meshcnn_b200 import dropin, mesh_prepare, synth
dropin.install(%(ref)r)
import models.networks as nets                                   # the reference's own networks.py, bound to the CUDA layers
from models.layers.mesh import Mesh                              # == meshcnn_b200.mesh.Mesh
from meshcnn_b200 import networks as N
E, B = 270, 3
datas = [mesh_prepare.from_arrays(*synth.make('ico', s, nu=3)) for s in range(B)]
x = np.stack([d.features for d in datas], 0)
x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
for arch, ncf, ncls in (('mconvnet', [16, 32], 4), ('meshunet', [16, 32, 64], 4)):
    opt = types.SimpleNamespace(norm='group', num_groups=4, pool_res=[200, 150], fc_n=10, resblocks=1)
    torch.manual_seed(0)
    theirs = nets.define_classifier(5, ncf, E, ncls, opt, [0], arch, 'normal', 0.02)      # DataParallel([0]) wrapper, on cuda:0
    mine = N.define_classifier(5, ncf, E, ncls, opt, [0], arch, 'normal', 0.02)
    mine.module.load_state_dict(theirs.module.state_dict())
    hold = arch == 'meshunet'
    y = torch.from_numpy(np.random.RandomState(1).randint(0, ncls, size=(B,) if not hold else (B, E))).cuda()
    res = []
    for net in (theirs, mine):
        net.train()
        meshes = np.array([Mesh(data=d, hold_history=hold) for d in datas], dtype=object)   # the loader hands an object array (F12)
        xin = torch.from_numpy(x).cuda().requires_grad_(True)
        out = net(xin, meshes)
        loss = torch.nn.functional.cross_entropy(out, y, ignore_index=-1)
        loss.backward()
        res.append((out.detach().cpu(), xin.grad.cpu(), {k: p.grad.cpu() for k, p in net.module.named_parameters()},
                    [m.edges_count for m in meshes]))
    (o1, g1, p1, e1), (o2, g2, p2, e2) = res
    def close(a, b, rtol, atol=1e-7):
        return float((a - b).abs().max()) <= rtol * float(b.abs().max()) + atol
    assert e1 == e2, (e1, e2)
    assert close(o1, o2, 1e-4), arch
    assert close(g1, g2, 1e-3), arch
    for k in p1:
        assert close(p1[k], p2[k], 1e-3, 2e-6), (arch, k)
# the fully-connected encoder head (networks.py:304-346): reference class vs ours, same weights
for gp in (None, 'avg', 'max'):
    torch.manual_seed(3)
    enc_t = nets.MeshEncoder([E, 200, 150], [5, 8, 16], fcs=[64, 10], blocks=1, global_pool=gp).cuda()
    enc_m = N.MeshEncoder([E, 200, 150], [5, 8, 16], fcs=[64, 10], blocks=1, global_pool=gp).cuda()
    enc_m.load_state_dict(enc_t.state_dict())
    outs = []
    for enc in (enc_t, enc_m):
        meshes = [Mesh(data=d, hold_history=True) for d in datas]
        code, skips = enc((torch.from_numpy(x).cuda(), meshes))
        outs.append(code.detach().cpu())
        assert code.shape == (B, 10) and len(skips) == 2
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-4 * float(outs[0].abs().max()) + 1e-6, gp
print('DROPIN-GPU-OK')
'''


def test_reference_networks_on_cuda_layers_match_own_networks():
    """reference networks.py (torch norm modules, [B, C, E, 1] tensors between layers) vs meshcnn_b200.networks (fused norm
    kernels, edge-major tensors) with the same weights: same output / gradients, same pooled sizes."""
    code = NETS % {'root': ROOT, 'ref': reference_arm.find_reference()}
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=900)
    assert 'DROPIN-GPU-OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.fixture()
def request_cleanup():
    import shutil
    paths = []
    yield paths
    for p in paths:
        shutil.rmtree(p, ignore_errors=True)


@pytest.mark.parametrize('mode', ['classification', 'segmentation'])
def test_reference_train_and_test_scripts_run_unchanged(tmp_path, request_cleanup, mode):
    """python -m meshcnn_b200.dropin <reference>/train.py ... : one epoch with data augmentation on a tiny synthetic dataset,
    checkpoint written; then <reference>/test.py loads it and reports an accuracy."""
    import shutil
    import tempfile
    from meshcnn_b200 import synth
    ref = reference_arm.find_reference()
    # the reference selects the files of a phase with `root.count(phase) == 1` (data/classification_data.py:56): the dataset
    # path must not contain the words "train" / "test" itself, which pytest's tmp_path does
    base = tempfile.mkdtemp(prefix='mcnn_ds_', dir='/tmp')
    assert 'train' not in base and 'test' not in base
    request_cleanup.append(base)
    data = os.path.join(base, 'data')
    ckpt = os.path.join(base, 'ckpt')
    if mode == 'classification':
        synth.write_classification_dataset(data, classes=2, n_train=4, n_test=2, nu=3)
        flags = ['--ncf', '16', '32', '--pool_res', '200', '150', '--norm', 'group', '--num_groups', '4', '--resblocks', '1',
                 '--ninput_edges', '270', '--fc_n', '10']
        aug = ['--num_aug', '3', '--flip_edges', '0.2', '--slide_verts', '0.2', '--scale_verts']
    else:
        synth.write_segmentation_dataset(data, n_train=4, n_test=2, nu=3, nclasses=3)
        flags = ['--dataset_mode', 'segmentation', '--arch', 'meshunet', '--ncf', '16', '32', '64', '--pool_res', '200', '150',
                 '--resblocks', '1', '--ninput_edges', '270']
        aug = ['--num_aug', '2', '--slide_verts', '0.2']
    common = ['--dataroot', data, '--name', 'tiny', '--checkpoints_dir', ckpt, '--batch_size', '2', '--gpu_ids', '0',
              '--num_threads', '0', '--seed', '0'] + flags
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', 'meshcnn_b200.dropin', os.path.join(ref, 'train.py')] + common + aug +
                       ['--niter', '1', '--niter_decay', '0', '--lr', '0.001', '--print_freq', '2', '--no_vis'],
                       capture_output=True, text=True, timeout=1200, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'End of epoch 1 / 1' in r.stdout and 'TEST ACC' in r.stdout, r.stdout[-3000:]
    assert '#training meshes = %d' % (8 if mode == 'classification' else 4) in r.stdout, r.stdout[-3000:]
    assert os.path.exists(os.path.join(ckpt, 'tiny', 'latest_net.pth'))
    sd = torch.load(os.path.join(ckpt, 'tiny', 'latest_net.pth'), map_location='cpu')
    assert any(k.endswith('conv.weight') and v.dim() == 4 and v.shape[-1] == 5 for k, v in sd.items())
    assert all(np.isfinite(v.numpy()).all() for v in sd.values() if v.is_floating_point())
    r = subprocess.run([sys.executable, '-m', 'meshcnn_b200.dropin', os.path.join(ref, 'test.py')] + common +
                       (['--export_folder', 'meshes'] if mode == 'segmentation' else []),
                       capture_output=True, text=True, timeout=1200, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'TEST ACC' in r.stdout, r.stdout[-3000:]
    if mode == 'segmentation':                                             # Mesh.export / export_segments through the lazy mirrors
        files = os.listdir(os.path.join(ckpt, 'tiny', 'meshes'))
        assert any(f.endswith('_0.obj') for f in files) and any(f.endswith('_2.obj') for f in files), files
