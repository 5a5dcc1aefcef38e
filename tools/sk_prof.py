"""Timeline of single CTAs of the stream-K forward kernel (needs `make prof`; MESHCNN_B200_LIB=.../libmeshcnn_b200_prof.so).
python tools/sk_prof.py E Cin Cout cta [cta ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch
dev = torch.device('cuda:0')
B = 64
E, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctas = [int(a) for a in sys.argv[4:]] or [0]
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev)
L = _lib.lib()
buf = (ctypes.c_longlong * 64)()
g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
ec = torch.full((B,), E, dtype=torch.int32, device=dev)
x = torch.randn(B, E, cin, device=dev)
w = torch.randn(cout, cin, 5, device=dev) * 0.05
out = torch.empty(B, E, cout, device=dev)
img = torch.empty(2 * 5 * cin * cout, device=dev)
ws = torch.zeros(int(L.mcnn_meshconv_fwd_tc_sk_ws()), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.mcnn_meshconv_tc_weight_images(w.data_ptr(), img.data_ptr(), None, cin, cout, st)
run = lambda: _lib.check(L.mcnn_meshconv_fwd_tc_sk(x.data_ptr(), g.data_ptr(), ec.data_ptr(), None, None, out.data_ptr(), img.data_ptr(), None, ws.data_ptr(), B, E, cin, cout, st), 'fwd')
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print('event-timed: %.1f us per launch (10 back to back)' % (a.elapsed_time(b) * 100))
for c in ctas:
    L.mcnn_debug_sk_prof(buf, c)
    run(); torch.cuda.synchronize()
    L.mcnn_debug_sk_prof(buf, c)
    t0 = buf[0]
    us = lambda i: (buf[i] - t0) / 1.965e3 if buf[i] else float('nan')
    print('CTA %d (E=%d %d->%d): entry %.1f, plan built %.1f, set-up done 0, producers done %.1f, TMEM released %.1f us' % (c, E, cin, cout, us(61), us(62), us(60), us(59)))
    for s in range(4):
        if not buf[1 + 4 * s]:
            break
        print('   seg %d: producers %.1f -> %.1f | last MMA issued %.1f | epilogue: enters %.1f, flags seen %.1f, accumulator ready %.1f, done %.1f'
              % (s, us(1 + 4 * s), us(2 + 4 * s), us(3 + 4 * s), us(23 + 4 * s), us(22 + 4 * s), us(20 + 4 * s), us(21 + 4 * s)))

import numpy as np
tb = (ctypes.c_ulonglong * 512)()
run(); run(); torch.cuda.synchronize()
L.mcnn_debug_sk_cta_times(tb)
t = np.array(list(tb), dtype=np.float64).reshape(256, 2)
t = t[t[:, 1] > 0]
t0 = t[:, 0].min()
st_, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
print('%d CTAs: entry %.1f .. %.1f us after the first, exit min %.1f  median %.1f  max %.1f us; duration min %.1f median %.1f max %.1f' % (len(t), st_.min(), st_.max(), en.min(), np.median(en), en.max(), (en - st_).min(), np.median(en - st_), (en - st_).max()))
order = np.argsort(en)
print('slowest CTAs (id: entry, exit):', ', '.join('%d: %.1f, %.1f' % (i, st_[i], en[i]) for i in order[-6:]))
