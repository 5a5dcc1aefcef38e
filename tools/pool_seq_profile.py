"""Per-stage clocks of the SEQUENTIAL collapse kernel (thread 0's stages and the helper warp's operations) on the cubes pool
chain.  Needs the diagnostics build:  make -C meshcnn_b200/csrc prof;
MESHCNN_B200_LIB=$PWD/meshcnn_b200/libmeshcnn_b200_prof.so python tools/pool_seq_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from meshcnn_b200.mesh import Mesh, MeshBatch
cfg = bench.CONFIGS['cubes']; dev = torch.device('cuda:0'); B, E = cfg['B'], cfg['E']
datas, x, labels = bench.make_inputs(cfg, B, 0)
meshes = [Mesh(data=d) for d in datas]
batch = MeshBatch(meshes, E, dev); batch.keep_clocks = True
torch.manual_seed(0)
net = bench.build_net(cfg, 'ours').to(dev).train()
with torch.no_grad(): net(torch.from_numpy(x).to(dev), meshes)
torch.cuda.synchronize(); batch.check()
nm = ['has_boundaries','clean_side','one_ring_request','pool_side x2','merge','post(in above)','wait_result']
for rec in batch.records:
    clk = rec.clocks.cpu().numpy().astype(np.float64)
    pops = clk[:,7]; col = (clk[:,3]-clk[:,2])
    print('pool %d->%d collapse phase %.1f us, pops %.1f' % (rec.E_in, rec.T, col.mean()/1965, pops.mean()))
    for i,s in zip((8,9,10,11,12,5,6), nm):
        print('   %-18s %.1f us total, %.0f cycles/pop' % (s, clk[:,i].mean()/1965, (clk[:,i]/pops).mean()))
    print('   helper: one_ring %.1f us, remove_scan %.1f us, merge %.1f us; inner %s' % (clk[:,13].mean()/1965, clk[:,14].mean()/1965, clk[:,15].mean()/1965, (clk[:,16:19].mean(0)/1965).round(1)))
