"""Cycle breakdown of CTA 0 of the tcgen05 forward kernel (needs `make prof`; MESHCNN_B200_LIB=.../libmeshcnn_b200_prof.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshcnn_b200 import _lib, synth, mesh_prepare
from meshcnn_b200.mesh import Mesh, MeshBatch
dev = torch.device('cuda:0')
B = 64
meshes = [Mesh(data=mesh_prepare.from_arrays(*synth.make('ico', s, nu=5))) for s in range(B)]
batch = MeshBatch(meshes, 750, dev)
L = _lib.lib()
buf = (ctypes.c_longlong * 16)()
for E, cin, cout in [(450, 256, 256), (600, 128, 128), (750, 64, 64)]:
    g = batch.level0.gemm[:, :E].clone().clamp_(max=E - 1).contiguous()
    ec = torch.full((B,), E, dtype=torch.int32, device=dev)
    x = torch.randn(B, E, cin, device=dev)
    w = torch.randn(cout, cin, 1, 5, device=dev) * 0.05
    out = torch.empty(B, E, cout, device=dev)
    wt = torch.empty(2 * 5 * cin, cout, device=dev)
    run = lambda: _lib.check(L.mcnn_meshconv_fwd_tc(x.data_ptr(), g.data_ptr(), ec.data_ptr(), w.data_ptr(), None, out.data_ptr(), wt.data_ptr(), B, E, cin, cout, torch.cuda.current_stream().cuda_stream), 'fwd')
    for _ in range(3): run()
    torch.cuda.synchronize()
    L.mcnn_debug_tc_prof(buf, 1)
    n = 5
    for _ in range(n): run()
    torch.cuda.synchronize()
    L.mcnn_debug_tc_prof(buf, 1)
    v = [x_ / n / 1.965e3 for x_ in buf]
    kb = buf[6] / n
    print('E=%d %d->%d  CTA 0, %d K blocks: producer thread 0: wait-empty %.1f us, split+store %.1f us, fence+arrive %.1f us | MMA warp wait-full %.1f us | epilogue wait for accumulator %.1f us'
          % (E, cin, cout, kb, v[0], v[1], v[2], v[3], v[4]))
