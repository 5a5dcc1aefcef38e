"""Data-gradient MeshConv kernel: persistent form vs one CTA per item, per layer shape of the cubes net (64 meshes).
python tools/bd_bench.py"""
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
    w = torch.randn(cout, cin, 5, device=dev) * 0.05
    img = torch.empty(2 * cin * 5 * cout, device=dev)
    dx = torch.zeros(B, E, cin, device=dev)
    signs = torch.randint(-1, 2, (2, B * E, cin), dtype=torch.int8, device=dev)
    gemm = lvl.gemm.view(B, 750, 4)[:, :E].clamp(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.mcnn_meshconv_tc_weight_images(w.data_ptr(), None, img.data_ptr(), cin, cout, st)
    call = lambda: L.mcnn_meshconv_bwd_data_tc_signs(dout.data_ptr(), None, signs.data_ptr(), gemm.data_ptr(), ec.data_ptr(), None, dx.data_ptr(), img.data_ptr(), B, E, cin, cout, st)
    L.mcnn_debug_set_tc_bd_tile(1)
    dx.zero_(); call(); r0 = dx.clone()
    t0 = timeit(call)
    L.mcnn_debug_set_tc_bd_tile(0)
    dx.zero_(); call(); r1 = dx.clone()
    t1 = timeit(call)
    err = float((r1 - r0).abs().max() / r0.abs().max())
    print('E %d %d->%d: items %d  tile kernel %.1f us, persistent %.1f us (%.2fx), rel diff %.1e' % (E, cin, cout, (B * E + 127) // 128 * (cin // 32), t0, t1, t0 / t1, err), flush=True)
    if len(sys.argv) > 1:
        res = []
        for knob in (0, 16, 2, 4, 1, 8, 20, 22, 31):
            L.mcnn_debug_set_tc(knob)
            res.append('%d: %.1f' % (knob, timeit(call)))
        L.mcnn_debug_set_tc(0)
        print('      persistent, knobs (1 no loads, 2 no MMA, 4 no st.shared, 8 no B copies, 16 no scatter, 32 no sign loads): ' + '  '.join(res), flush=True)
    if os.environ.get('BD_PROF'):
        import ctypes
        buf = (ctypes.c_longlong * 64)()
        for knob in (0, 31, 16):
            L.mcnn_debug_set_tc(knob)
            call(); torch.cuda.synchronize()
            L.mcnn_debug_sk_prof(buf, 70)
            call(); torch.cuda.synchronize()
            L.mcnn_debug_sk_prof(buf, 70)
            v = [b_ / 1.965e3 for b_ in buf]
            print('      knob %d, CTA 70 (us at 1965 MHz): %d steps | producer thread 0: total %.1f, wait-empty %.1f, fence+arrive %.1f | MMA warp: total %.1f, wait-full %.1f, wait-acc-empty %.1f | epilogue thread: total %.1f, wait-acc-full %.1f | elected thread: MMA issue %.1f, commits %.1f'
                  % (knob, buf[3], v[0], v[1], v[2], v[4], v[6], v[5], v[7], v[8], v[10], v[9]), flush=True)
        L.mcnn_debug_set_tc(0)
