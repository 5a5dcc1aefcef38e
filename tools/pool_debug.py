import faulthandler, os, sys
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from meshcnn_b200.mesh import Mesh, MeshBatch
from meshcnn_b200.mesh_pool import MeshPool
cfg = bench.CONFIGS['cubes']
dev = torch.device('cuda:0')
B, E = int(sys.argv[1]), cfg['E']
clocks = int(sys.argv[2])
datas, x, labels = bench.make_inputs(cfg, B, 0)
meshes = [Mesh(data=d) for d in datas]
batch = MeshBatch(meshes, E, dev)
batch.keep_clocks = bool(clocks)
net = bench.build_net(cfg, 'ours').to(dev).train()
xd = torch.from_numpy(x).to(dev)
print('start', B, clocks, flush=True)
with torch.no_grad():
    net(xd, meshes)
torch.cuda.synchronize()
batch.check()
print('ok', [int(r.status.sum()) for r in batch.records], flush=True)
