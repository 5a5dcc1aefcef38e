"""Sequential vs round collapse kernel by mesh size: one pool level (E -> 0.75 E) on a batch of tori, CUDA-event time of
MeshBatch.pool per kernel form.  python tools/pool_size_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshcnn_b200 import _lib, mesh_prepare, synth
from meshcnn_b200.mesh import Mesh, MeshBatch

dev = torch.device('cuda:0')
L = _lib.lib()
for m, n in ((20, 38), (30, 45), (34, 49), (44, 53), (50, 60)):
    B = 16
    datas = []
    for s in range(B):
        vs, faces = synth.make('torus', s, m=m, n=n)
        datas.append(mesh_prepare.from_arrays(vs, faces, filename='t%d.obj' % s))
    E = max(d.edges_count for d in datas)
    rs = np.random.RandomState(0)
    norms = torch.from_numpy(rs.rand(B, E).astype(np.float32)).to(dev)
    res = {}
    for mode in (2, 1):
        L.mcnn_debug_set_pool_mode(mode)
        ts = []
        for rep in range(4):
            batch = MeshBatch([Mesh(data=d) for d in datas], E, dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            batch.pool(norms, int(0.75 * E) // 2 * 2)
            e1.record()
            torch.cuda.synchronize()
            batch.check()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[mode] = min(ts[1:])
    L.mcnn_debug_set_pool_mode(0)
    print('E = %5d: sequential %.0f us, rounds %.0f us' % (E, res[2], res[1]))
