"""Weight-gradient MeshConv kernel per layer shape of the cubes net (64 meshes), with the attribution knobs.
python tools/bw_bench.py [knobs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch

dev = torch.device('cuda:0')
L = _lib.lib()
B = 64
bt = MeshBatch([Mesh(data=mesh_prepare.from_arrays(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25))))) for s in range(B)], 750, dev)
lvl = bt.level(750)


def timeit(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3


for E, cin, cout in [(750, 64, 64), (600, 64, 128), (600, 128, 128), (450, 128, 256), (450, 256, 256), (300, 256, 256)]:
    dout = torch.randn(B, E, cout, device=dev)
    x = torch.randn(B, E, cin, device=dev)
    dw = torch.empty(cout, cin, 5, device=dev)
    db = torch.empty(cout, device=dev)
    ws = torch.empty(int(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, cin, cout)), device=dev)
    gemm = lvl.gemm.view(B, 750, 4)[:, :E].clamp(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: L.mcnn_meshconv_bwd_weight_tc(dout.data_ptr(), x.data_ptr(), gemm.data_ptr(), ec.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), B, E, cin, cout, st)
    res = []
    for on in (0, 1):                                   # first-generation kernel (both operands in shared memory) vs dout^T in TMEM
        L.mcnn_debug_set_tc_bw_tmem_a(on)
        res.append('%s: %.1f' % ('tmem-A' if on else 'smem-A', timeit(call)))
    if cout % 256 == 0:
        for ot in (1, 2):
            L.mcnn_debug_set_tc_bw_ot(ot)
            ws = torch.empty(int(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, cin, cout)), device=dev)
            res.append('tiles/CTA %d: %.1f' % (ot, timeit(call)))
        L.mcnn_debug_set_tc_bw_ot(0)
        ws = torch.empty(int(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, cin, cout)), device=dev)
    for knob in ((0, 1, 4, 5, 2, 7) if len(sys.argv) > 1 else (0,)):
        L.mcnn_debug_set_tc(knob)
        res.append('%d: %.1f' % (knob, timeit(call)))
    L.mcnn_debug_set_tc(0)
    print('E %d %d->%d: us (knobs 1 no loads, 2 no MMA, 4 no st.shared; includes the split-reduce and bias kernels): %s' % (E, cin, cout, '  '.join(res)), flush=True)
