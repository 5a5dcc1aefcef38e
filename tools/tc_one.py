"""Run one tensor-core MeshConv forward shape a few times (for ncu).  python tools/tc_one.py E Cin Cout [pair]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch
E, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pair = len(sys.argv) > 4 and sys.argv[4] == 'pair'
dev = torch.device('cuda:0')
B = 64
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev)
L = _lib.lib()
L.mcnn_debug_set_tc_pair(1 if pair else 0)
g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
ec = torch.full((B,), E, dtype=torch.int32, device=dev)
x = torch.randn(B, E, cin, device=dev)
w = torch.randn(cout, cin, 1, 5, device=dev) * 0.05
out = torch.empty(B, E, cout, device=dev)
wt = torch.empty(2 * 5 * cin, cout, device=dev)
for _ in range(4):
    _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), None, out.data_ptr(), wt.data_ptr(), B, E, cin, cout, torch.cuda.current_stream().cuda_stream), 'fwd')
torch.cuda.synchronize()
print('ok')
