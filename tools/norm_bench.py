"""GroupNorm block forward / backward: one-launch form vs statistics / finalize / apply, per layer shape of the cubes net.
python tools/norm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from meshcnn_b200 import _lib
from meshcnn_b200.functional import fused_norm

dev = torch.device('cuda:0')
L = _lib.lib()


def timeit(fn, k=30):
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


shapes = [(64, 750, 64, 'group'), (64, 600, 128, 'group'), (64, 450, 256, 'group'), (64, 300, 256, 'group'), (12, 2280, 32, 'instance'), (12, 1800, 64, 'instance')]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) if v.isdigit() else v for v in a.split(',')) for a in sys.argv[1:]]
for B, E, C, kind in shapes:
    norm = (nn.GroupNorm(16, C) if kind == 'group' else nn.InstanceNorm2d(C)).to(dev).train()
    x = torch.randn(B, E, C, device=dev, requires_grad=True)
    aux = torch.randn(B, E, C, device=dev, requires_grad=True)
    dy = torch.randn(B, E, C, device=dev)
    out = []
    for on in (0, 1):
        L.mcnn_debug_set_norm_fused(on)
        g = torch.cuda.CUDAGraph()          # graph replay: no host launch overhead in the comparison
        y = fused_norm(x, norm, pre_add=aux, pre_relu=True, post_relu=True)
        y.backward(dy)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                y = fused_norm(x, norm, pre_add=aux, pre_relu=True, post_relu=True)
                y.backward(dy)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            y = fused_norm(x, norm, pre_add=aux, pre_relu=True, post_relu=True)
            y.backward(dy)
        out.append(timeit(g.replay))
    L.mcnn_debug_set_norm_fused(1)
    print('B %d E %d C %d %s: forward + backward  three kernels %.1f us, one launch %.1f us (%.2fx)' % (B, E, C, kind, out[0], out[1], out[0] / out[1]), flush=True)
