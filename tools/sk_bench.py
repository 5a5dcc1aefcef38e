"""Forward MeshConv: persistent stream-K kernel vs one CTA per tile, per layer shape of the cubes net (64 meshes).
python tools/sk_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from meshcnn_b200 import _lib, synth
from meshcnn_b200 import functional as Fn
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200 import mesh_prepare

dev = torch.device('cuda:0')
L = _lib.lib()
if os.environ.get('SK_NTILE'):
    L.mcnn_debug_set_tc_ntile(int(os.environ['SK_NTILE']))
B = 64
datas = [mesh_prepare.from_arrays(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]


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


bt = MeshBatch([Mesh(data=d) for d in datas], 750, dev)
lvl = bt.level(750)
ws = torch.zeros(int(L.mcnn_meshconv_fwd_tc_sk_ws()), dtype=torch.uint8, device=dev)
SHAPES = [(750, 64, 64), (600, 64, 128), (600, 128, 128), (450, 128, 256), (450, 256, 256), (300, 256, 256)]
if os.environ.get('SK_SHAPES'):
    SHAPES = [tuple(int(v) for v in t.split(',')) for t in os.environ['SK_SHAPES'].split(';')]
for E, cin, cout in SHAPES:
    # the topology of the 750-edge level serves every E (gather ids < E are what matters for timing): use the first E edges
    x = torch.randn(B, E, cin, device=dev)
    w = torch.randn(cout, cin, 5, device=dev) * 0.05
    bias = torch.randn(cout, device=dev)
    img = torch.empty(2 * cin * 5 * cout, device=dev)
    out = torch.empty(B, E, cout, device=dev)
    gemm = lvl.gemm.view(B, 750, 4)[:, :E].clamp(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.mcnn_meshconv_tc_weight_images(w.data_ptr(), img.data_ptr(), None, cin, cout, st)
    t0 = timeit(lambda: L.mcnn_meshconv_fwd_tc_signs(x.data_ptr(), gemm.data_ptr(), ec.data_ptr(), None, bias.data_ptr(), out.data_ptr(), img.data_ptr(), None, B, E, cin, cout, st))
    o0 = out.clone()
    t1 = timeit(lambda: L.mcnn_meshconv_fwd_tc_sk(x.data_ptr(), gemm.data_ptr(), ec.data_ptr(), None, bias.data_ptr(), out.data_ptr(), img.data_ptr(), None, ws.data_ptr(), B, E, cin, cout, st))
    err = float((out - o0).abs().max() / o0.abs().max())
    print('E %d %d->%d: tiles %d  tile kernel %.1f us, stream-K %.1f us (%.2fx), rel diff %.1e' % (E, cin, cout, (B * E + 127) // 128 * max(1, cout // 256), t0, t1, t0 / t1, err), flush=True)
    if len(sys.argv) > 1:
        for gsz in [int(a) for a in sys.argv[1:]]:
            L.mcnn_debug_set_tc_sk_ctas(gsz)
            t = timeit(lambda: L.mcnn_meshconv_fwd_tc_sk(x.data_ptr(), gemm.data_ptr(), ec.data_ptr(), None, bias.data_ptr(), out.data_ptr(), img.data_ptr(), None, ws.data_ptr(), B, E, cin, cout, st))
            print('      grid %d: %.1f us' % (gsz, t), flush=True)
        L.mcnn_debug_set_tc_sk_ctas(0)
