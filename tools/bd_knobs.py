"""Attribution for the tcgen05 dgrad kernel: normal / no sign loads / no scatter at all (results wrong when a knob is set)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch
dev = torch.device('cuda:0'); B = 64
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev); L = _lib.lib()
for E, cin, cout in [(750, 64, 64), (600, 128, 128), (450, 256, 256), (300, 256, 256)]:
    g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    x = torch.randn(B, E, cin, device=dev); w = torch.randn(cout, cin, 1, 5, device=dev) * 0.05
    dout = torch.randn(B, E, cout, device=dev); dx = torch.zeros(B, E, cin, device=dev)
    wt = torch.empty(2 * 5 * cin, cout, device=dev)
    res = []
    for knob in (0, 32, 16):
        L.mcnn_debug_set_tc(knob)
        run = lambda: _lib.check(L.mcnn_meshconv_bwd_data_tc(dout.data_ptr(), x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), dx.data_ptr(), wt.data_ptr(), B, E, cin, cout, torch.cuda.current_stream().cuda_stream), 'bd')
        for _ in range(3): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) * 100)
    L.mcnn_debug_set_tc(0)
    print('E=%d %d->%d  dgrad us: normal %.1f | no sign loads %.1f | no scatter (GEMM + set-up only) %.1f' % (E, cin, cout, *res))
