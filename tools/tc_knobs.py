"""Bottleneck attribution for the tcgen05 MeshConv forward kernel: time each bench-config layer with the debug knobs
(skip global loads / skip MMA issue / skip st.shared).  Results are wrong by construction when a knob is set."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200.mesh_conv import MeshConv

dev = torch.device('cuda:0')
B = 64
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev)
L = _lib.lib()
shapes = [(750, 64, 64), (600, 64, 128), (600, 128, 128), (450, 128, 256), (450, 256, 256), (300, 256, 256)]
for E, cin, cout in shapes:
    ms = meshes
    conv = MeshConv(cin, cout, bias=False).to(dev)
    x = torch.randn(B, E, cin, device=dev)
    lv = batch.level0 if E == 750 else None
    # a level with E rows: reuse level 0 tables truncated (neighbour ids >= E are clamped to keep the gather in range)
    g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
    from meshcnn_b200.mesh import Level
    level = Level(E, g, batch.level0.sides[:, :E].contiguous(), torch.full((B,), E, dtype=torch.int32, device=dev))
    out = torch.empty(B, E, cout, device=dev)
    wt = torch.empty(2 * 5 * cin, cout, device=dev)
    res = []
    for knob in (0, 1, 2, 4, 3, 7, 15, 8, 9):
        L.mcnn_debug_set_tc(knob)
        torch.cuda.synchronize()
        def run():
            _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(), conv.conv.weight.data_ptr(), None,
                                              out.data_ptr(), wt.data_ptr(), B, E, cin, cout, torch.cuda.current_stream().cuda_stream), 'fwd')
        for _ in range(3): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 10 * 1e3)
    L.mcnn_debug_set_tc(0)
    flops = 2.0 * B * E * 5 * cin * cout
    print('E=%d %d->%d  us: normal %.1f | noload %.1f | nomma %.1f | nosts %.1f | noload+nomma %.1f | all-off %.1f | all-off, no B copies %.1f | only B copies off %.1f | B copies + gathers off %.1f  (%.1f TFLOP/s fp32-equivalent)'
          % (E, cin, cout, *res, flops / res[0] / 1e6))
