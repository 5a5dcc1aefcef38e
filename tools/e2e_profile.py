"""cProfile of bench.py's end-to-end step (host buffers in, loss out) to find host-side stalls."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from meshcnn_b200 import ddp
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200.optim import FlatAdam

cfg = bench.CONFIGS['cubes']
dev = torch.device('cuda:0')
B, E = cfg['B'], cfg['E']
datas, x, labels = bench.make_inputs(cfg, B, 0)
net = bench.build_net(cfg, 'ours').to(dev).train()
reducer = ddp.FlatGradReducer(net)
opt = FlatAdam(reducer, lr=cfg['lr'], betas=(0.9, 0.999))
crit = bench.loss_fn_for(cfg)
x_pin = torch.from_numpy(x).pin_memory()
y_pin = torch.from_numpy(labels).pin_memory()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fresh = [[Mesh(data=d) for d in datas] for _ in range(K + 2)]


def step(ms):
    bt = MeshBatch(ms, E, dev)
    xin = x_pin.to(dev, non_blocking=True).requires_grad_(True)
    yin = y_pin.to(dev, non_blocking=True)
    reducer.zero()
    out = net(xin, ms)
    loss = crit(out, yin)
    loss.backward()
    reducer.allreduce()
    opt.step()
    return float(loss.item())


for i in range(2):
    step(fresh[i])
torch.cuda.synchronize()
ts = []
pr = cProfile.Profile()
pr.enable()
for i in range(K):
    t0 = time.perf_counter()
    step(fresh[2 + i])
    ts.append((time.perf_counter() - t0) * 1e3)
pr.disable()
print('per-step ms:', ' '.join('%.1f' % t for t in ts))
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
