"""What does the gradient all-reduce cost INSIDE a replayed step graph (after ~1 ms of compute, ranks arriving together),
as opposed to back-to-back (tools/allreduce_probe.py)?  NCCL vs the symmetric-memory all-reduces of torch.
torchrun --nproc-per-node N tools/ar_in_step_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1322578
a = torch.randn(4096, 4096, device=dev)
buf = torch.randn(n, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def compute():
    x = a
    for _ in range(6):
        x = x @ a * 1e-3
    return x


def make_graph(ar):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            compute(); ar(); buf.mul_(1.0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        compute()
        ar()
        buf.mul_(1.0)
    return g


def timeit(g, k=30):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(k):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = torch.tensor([sum(x.elapsed_time(y) for x, y in evs) / k], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


res = {}
res['no collective'] = timeit(make_graph(lambda: None))
res['NCCL all_reduce (SUM)'] = timeit(make_graph(lambda: dist.all_reduce(buf)))
res['NCCL all_reduce (AVG)'] = timeit(make_graph(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG)))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(n, dtype=torch.float32, device=dev)
    symm.rendezvous(t, dist.group.WORLD.group_name)
    t.copy_(buf)
    for name in ('one_shot_all_reduce', 'two_shot_all_reduce_', 'multimem_all_reduce_'):
        try:
            op = getattr(torch.ops.symm_mem, name)
            res['symm_mem.' + name] = timeit(make_graph(lambda op=op: op(t, 'sum', dist.group.WORLD.group_name)))
        except Exception as e:      # noqa: BLE001
            res['symm_mem.' + name] = str(e)[:80]
except Exception as e:              # noqa: BLE001
    res['symm_mem'] = 'unavailable: %s' % str(e)[:80]
if rank == 0:
    base = res['no collective']
    print('world %d, %d floats; graph = 6 matmuls + all-reduce + 1 elementwise; us per replay (max over ranks):' % (dist.get_world_size(), n))
    for k, v in res.items():
        print('  %-32s %s' % (k, ('%.1f  (+%.1f)' % (v, v - base)) if isinstance(v, float) else v), flush=True)
dist.destroy_process_group()
