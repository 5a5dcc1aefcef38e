"""World-size-2 gloo test of the data-parallel plumbing (host logic only; the kernels are not involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meshcnn_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)                      # different initial weights per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ddp.broadcast_parameters(net)
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()])
    B = 7                                        # uneven: 4 + 3
    lo, hi = ddp.shard_range(B, rank, world)
    red = ddp.FlatGradReducer(net, local_items=hi - lo, global_items=B)
    torch.manual_seed(123)
    x, y = torch.randn(B, 6), torch.randint(0, 3, (B,))
    red.zero()
    loss = torch.nn.functional.cross_entropy(net(x[lo:hi]), y[lo:hi])     # mean over the local shard
    loss.backward()
    assert red.check_views()
    red.allreduce()
    # single-process truth: mean over the whole batch
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    torch.nn.functional.cross_entropy(ref(x), y).backward()
    g_ref = torch.cat([p.grad.flatten() for p in ref.parameters()])
    q.put((rank, lo, hi, float((red.flat - g_ref).abs().max()), w0.sum().item()))
    dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]
    assert all(r[3] < 1e-6 for r in res), res                 # weighted mean of shard gradients == full-batch gradient
    assert abs(res[0][4] - res[1][4]) < 1e-7                   # parameters were broadcast from rank 0


def test_shard_range_covers_batch():
    for B in (1, 7, 12, 64):
        for w in (1, 2, 4, 8):
            spans = [ddp.shard_range(B, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
