"""Golden vectors for Mesh.export / export_segments (reference mesh.py:74-124): the UNMODIFIED reference pools a small
mesh twice with --export_folder semantics; the text of every file it writes is stored.  Build container only."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from ref_loader import load_reference      # noqa: E402
from meshcnn_b200 import mesh_prepare, synth   # noqa: E402
import types                               # noqa: E402

ref = load_reference()
tmp = tempfile.mkdtemp(prefix='golden_export_')
vs, faces = synth.make('ico', 5, nu=3)
obj = os.path.join(tmp, 'ico270.obj')
synth.write_obj(obj, vs, faces)
vs = mesh_prepare.read_obj(obj)[0]
folder = os.path.join(tmp, 'exp')
os.makedirs(folder)
mesh = ref.Mesh(file=obj, opt=types.SimpleNamespace(num_aug=1), hold_history=True, export_folder=folder)
targets = [200, 150]
torch.manual_seed(7)
d = {'vs0': vs, 'faces': faces, 'targets': np.array(targets), 'name': 'ico270.obj'}
E = mesh.edges_count
for li, T in enumerate(targets):
    x = torch.randn(1, 6, E)
    d['P%d_norms' % li] = torch.sum(x[0] * x[0], 0).numpy().astype(np.float32)
    out = ref.MeshPool(T)(x, [mesh])
    E = T
rs = np.random.RandomState(1)
seg = rs.randint(0, 4, size=270)
d['segments'] = seg
mesh.export_segments(seg)
for i in range(len(targets) + 1):
    d['file%d' % i] = np.frombuffer(open(os.path.join(folder, 'ico270_%d.obj' % i), 'rb').read(), dtype=np.uint8)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'export_ico270.npz'), **d)
print('wrote export_ico270.npz', [len(d['file%d' % i]) for i in range(3)])
