"""Cost of the step's one collective in isolation: all-reduce of the flat fp32 gradient (1.32 M floats for the cubes net),
launched eagerly and replayed from a CUDA graph.  torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1322000
buf = torch.randn(n, device=dev)
for _ in range(5):
    dist.all_reduce(buf)
torch.cuda.synchronize()
dist.barrier()


def timeit(fn, k=50):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3


eager = timeit(lambda: dist.all_reduce(buf))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dist.all_reduce(buf)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    dist.all_reduce(buf)
graph = timeit(g.replay)
res = ''
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(n, dtype=torch.float32, device=dev)
    symm.rendezvous(t, dist.group.WORLD.group_name)
    t.copy_(buf)
    for name in ('one_shot_all_reduce', 'two_shot_all_reduce_', 'multimem_all_reduce_'):
        try:
            op = getattr(torch.ops.symm_mem, name)
            f = (lambda op=op: op(t, 'sum', dist.group.WORLD.group_name))
            f()
            torch.cuda.synchronize()
            res += ' %s %.1f us' % (name, timeit(f))
        except Exception as e:      # noqa: BLE001
            res += ' %s: %s' % (name, str(e)[:60])
except Exception as e:              # noqa: BLE001
    res = 'symm_mem unavailable: %s' % str(e)[:80]
if rank == 0:
    print('world %d, %d floats: NCCL eager %.1f us, NCCL in graph %.1f us | %s' % (dist.get_world_size(), n, eager, graph, res), flush=True)
os._exit(0)
