"""Phase breakdown of the MeshPool collapse kernel (SM clocks at the phase boundaries of every mesh's CTA) on the pool
chain of a bench config.  python tools/pool_phases.py [cubes|humanseg]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200.mesh_pool import MeshPool

name = sys.argv[1] if len(sys.argv) > 1 else 'cubes'
cfg = bench.CONFIGS[name]
dev = torch.device('cuda:0')
B, E = cfg['B'], cfg['E']
datas, x, labels = bench.make_inputs(cfg, B, 0)
meshes = [Mesh(data=d) for d in datas]
batch = MeshBatch(meshes, E, dev)
batch.keep_clocks = True
torch.manual_seed(0)
names = ['load', 'sort', 'collapse', 'clean + groups export']
net = bench.build_net(cfg, 'ours').to(dev).train()
xd = torch.from_numpy(x).to(dev)
with torch.no_grad():
    net(xd, meshes)
torch.cuda.synchronize()
batch.check()
for rec in batch.records:
    clk = rec.clocks.cpu().numpy().astype(np.float64)
    d = np.diff(clk[:, :5], axis=1) / 1.965e3            # us at 1965 MHz
    tot = (clk[:, 4] - clk[:, 0]) / 1.965e3
    print('%s pool %d -> %d: CTA total mean %.1f max %.1f us' % (name, rec.E_in, rec.T, tot.mean(), tot.max()))
    for i, nm in enumerate(names):
        print('    %-14s mean %7.1f  max %7.1f us' % (nm, d[:, i].mean(), d[:, i].max()))
    print('    pops mean %.1f max %d; removed edges %d' % (clk[:, 7].mean(), clk[:, 7].max(), rec.E_in - rec.T))
    rounds = clk[:, 8]
    print('    rounds mean %.1f max %d; pops per round %.2f; collapse phase per round %.2f us' % (
        rounds.mean(), rounds.max(), clk[:, 7].sum() / max(rounds.sum(), 1), (d[:, 2] / np.maximum(rounds, 1)).mean()))
    print('    per round (us): select %.2f | inspect + reserve %.2f | check %.2f | commit %.2f;  rounds that ran one candidate alone: %.1f of %.1f; committed per round %.2f' % (
        tuple((clk[:, i] / np.maximum(rounds, 1)).mean() / 1.965e3 for i in (9, 10, 11, 12)) + ((clk[:, 13] // 1000000).mean(), rounds.mean(), (clk[:, 14] // 1000000).sum() / max(rounds.sum(), 1))))
    print('      warp 0: one-ring walks %d, of them on single-slab lists %d; COLLAPSE verdicts %d, of them with nine distinct rows %d' % (
        ((clk[:, 13] // 1000) % 1000).sum(), (clk[:, 13] % 1000).sum(), ((clk[:, 14] // 1000) % 1000).sum(), (clk[:, 14] % 1000).sum()))
    print('      warp 0, per round (us): inspect = boundary/side tests %.2f + one-ring %.2f (when reached), own work in total %.2f, rest = wait at the barrier;  commit = prefix %.2f + collapse %.2f + re-pack/log %.2f, rest = wait' % tuple(
        (clk[:, i] / np.maximum(rounds, 1)).mean() / 1.965e3 for i in (15, 16, 17, 18, 5, 6)))
    e = [clk[:, 3]] + [clk[:, i] for i in (19, 20, 21, 22, 23)] + [clk[:, 4]]
    print('    . export: E1+adjacency %.1f | E2 %.1f | E3 %.1f | occ %.1f | BFS columns %.1f | CSC/CSR write %.1f us (mean);  max %s' % (
        tuple(((e[i + 1] - e[i]) / 1.965e3).mean() for i in range(6)) + (' '.join('%.0f' % ((e[i + 1] - e[i]) / 1.965e3).max() for i in range(6)),)))
