"""Time the tcgen05 MeshConv kernels of the bench layers for every admissible forward N tile (and the backward kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch, Level

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev)
L = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


shapes = [(750, 64, 64), (600, 64, 128), (600, 128, 128), (450, 128, 256), (450, 256, 256), (300, 256, 256)]
for E, cin, cout in shapes:
    g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    x = torch.randn(B, E, cin, device=dev)
    w = torch.randn(cout, cin, 1, 5, device=dev) * 0.05
    out = torch.empty(B, E, cout, device=dev)
    dout = torch.randn(B, E, cout, device=dev)
    dx = torch.zeros(B, E, cin, device=dev)
    dw = torch.empty_like(w)
    wt = torch.empty(2 * 5 * cin, cout, device=dev)
    ws = torch.empty(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, cin, cout), device=dev)
    res = []
    for nt in (64, 128, 256):
        if cout % nt:
            res.append(float('nan')); continue
        L.mcnn_debug_set_tc_ntile(nt)
        res.append(timeit(lambda: _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), None, out.data_ptr(), wt.data_ptr(), B, E, cin, cout, st()), 'fwd')))
    L.mcnn_debug_set_tc_ntile(0)
    pair = float('nan')
    L.mcnn_debug_set_tc_pair(2)
    pw16 = timeit(lambda: _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), None, out.data_ptr(), wt.data_ptr(), B, E, cin, cout, st()), 'fwd'))
    L.mcnn_debug_set_tc_pair(0)
    if cout % 256 == 0:
        L.mcnn_debug_set_tc_pair(1)
        pair = timeit(lambda: _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), None, out.data_ptr(), wt.data_ptr(), B, E, cin, cout, st()), 'fwd'))
        L.mcnn_debug_set_tc_pair(0)
    bd = timeit(lambda: _lib.check(L.mcnn_meshconv_bwd_data_tc(dout.data_ptr(), x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), dx.data_ptr(), wt.data_ptr(), B, E, cin, cout, st()), 'bd'))
    bw = timeit(lambda: _lib.check(L.mcnn_meshconv_bwd_weight_tc(dout.data_ptr(), x.data_ptr(), g.data_ptr(), ec.data_ptr(), dw.data_ptr(), None, ws.data_ptr(), B, E, cin, cout, st()), 'bw'))
    print('E=%d %d->%d  fwd us: N64 %.1f | N128 %.1f | N256 %.1f | pair %.1f | 16 producer warps %.1f   bwd_data %.1f  bwd_weight %.1f   (m-tiles %d)' % (E, cin, cout, *res, pair, pw16, bd, bw, (B * E + 127) // 128))
