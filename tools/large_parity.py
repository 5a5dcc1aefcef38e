"""One-off full-size parity check for BASELINE config 5: a 169 932-edge torus pooled 170 000 -> 128 000 -> 96 000 by the
workspace variant of the collapse kernel against the CPU oracle (sparse groups) driven with the same fp32 norms.
Bit-exact: pop sequence, outcomes, pooled topology, occurrences.  Takes a few minutes of CPU time (the oracle)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from meshcnn_b200 import mesh_prepare, synth
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200.mesh_pool import MeshPool
from oracle import meshcnn_oracle as O

dev = torch.device('cuda:0')
raw = synth.make('torus', 0, m=238, n=238)
data = mesh_prepare.from_arrays(*raw)
mesh = Mesh(data=data)
E0 = 170000
batch = MeshBatch([mesh], E0, dev)
batch.keep_logs = True
om = O.OracleMesh(mesh_prepare.from_arrays(*raw))
torch.manual_seed(0)
x = torch.zeros(1, 8, E0, device=dev)
x[:, :, :mesh.edges_count] = torch.randn(1, 8, mesh.edges_count, device=dev)
for T in (128000, 96000):
    pool = MeshPool(T)
    t0 = time.time()
    out = pool(x, [mesh])
    torch.cuda.synchronize()
    tg = time.time() - t0
    batch.check()
    rec = pool.last_record
    norms = rec.norms.cpu().numpy()
    t0 = time.time()
    o_out, log = O.pool_mesh(om, x[0].cpu(), T, norms=norms[0])
    tc = time.time() - t0
    npops = int(rec.logs['npops'][0])
    np.testing.assert_array_equal(rec.logs['pops'][0, :npops].cpu().numpy(), np.asarray(log['pops']))
    np.testing.assert_array_equal(rec.logs['outcome'][0, :npops].cpu().numpy(), np.asarray(log['outcomes']))
    assert mesh.edges_count == om.edges_count
    np.testing.assert_array_equal(mesh.gemm_edges, om.gemm_edges)
    np.testing.assert_array_equal(mesh.sides, om.sides)
    np.testing.assert_array_equal(mesh.edges, om.edges)
    np.testing.assert_array_equal(rec.occ[0, :len(log['occurrences'])].cpu().numpy(), np.asarray(log['occurrences'], dtype=np.float32))
    err = float((out[0].cpu() - o_out).abs().max() / o_out.abs().max())
    print('pool -> %d: %d pops, %d edges left; GPU %.3f s, CPU oracle %.1f s (x%.0f); bit-exact topology, features rel err %.1e'
          % (T, npops, mesh.edges_count, tg, tc, tc / tg, err), flush=True)
    x = torch.tanh(out.detach()) + 0.05 * torch.randn_like(out)
print('large parity ok')
