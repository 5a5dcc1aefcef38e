"""Two ranks, two GPUs, NCCL: the all-reduced flat gradient of the sharded batch must equal the gradient one process computes
over the whole batch.  Uneven shards (3 meshes = 2 + 1) exercise the mesh-count weighting.  Skipped on a box
with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net_and_data():
    from meshcnn_b200 import mesh_prepare, synth
    from meshcnn_b200 import networks as N
    torch.manual_seed(0)
    E, B = 270, 3
    net = N.MeshEncoderDecoder([E, 200, 150], [5, 16, 32, 64], [64, 32, 16, 4], blocks=1, transfer_data=True)   # InstanceNorm:
    N.init_weights(net, 'normal', 0.02)                                        # no statistics across meshes, shards are exact
    datas = [mesh_prepare.from_arrays(*synth.make('ico', s, nu=3)) for s in range(B)]
    x = np.stack([d.features for d in datas], 0)
    x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
    y = np.random.RandomState(0).randint(0, 4, size=(B, E)).astype(np.int64)
    return net, datas, x, y, E, B


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from meshcnn_b200 import ddp
    from meshcnn_b200.mesh import Mesh
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    net, datas, x, y, E, B = _net_and_data()
    net = net.to(dev).train()
    ddp.broadcast_parameters(net)
    lo, hi = ddp.shard_range(B, rank, world)
    red = ddp.FlatGradReducer(net, local_items=hi - lo, global_items=B)
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1)
    red.zero()
    out = net(torch.from_numpy(x[lo:hi]).to(dev), [Mesh(data=d, hold_history=True) for d in datas[lo:hi]])
    crit(out, torch.from_numpy(y[lo:hi]).to(dev)).backward()
    assert red.check_views()
    red.allreduce()
    torch.cuda.synchronize()
    q.put((rank, lo, hi, red.flat.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_nccl_allreduced_gradient_equals_single_process_gradient():
    import torch.multiprocessing as mp
    from meshcnn_b200 import ddp
    from meshcnn_b200.mesh import Mesh
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 2), (2, 3)]
    np.testing.assert_array_equal(res[0][3], res[1][3])                        # every rank holds the same reduced gradient
    # single-process truth on GPU 0 (a): the same shards, one after the other, combined with the mesh-count weights -- the
    # same kernels on the same shapes as the ranks ran, so only the order of the atomic scatter-adds and the final fp32 sum
    # differ: 2e-5;  (b) the whole batch in one forward / backward (CE mean over all, equally many, edges): other split sizes
    # in the weight-gradient kernels and other summation orders throughout -- the fp32 gradient noise of a network
    # (tests/test_networks_gpu.py): 2e-3
    dev = torch.device('cuda', 0)

    def grad_of(lo, hi):
        net, datas, x, y, E, B = _net_and_data()
        net = net.to(dev).train()
        red = ddp.FlatGradReducer(net)
        red.zero()
        out = net(torch.from_numpy(x[lo:hi]).to(dev), [Mesh(data=d, hold_history=True) for d in datas[lo:hi]])
        torch.nn.CrossEntropyLoss(ignore_index=-1)(out, torch.from_numpy(y[lo:hi]).to(dev)).backward()
        return red.flat.double().cpu().numpy()

    sharded = (2.0 / 3.0) * grad_of(0, 2) + (1.0 / 3.0) * grad_of(2, 3)
    whole = grad_of(0, 3)
    scale = float(np.abs(whole).max())
    got = res[0][3].astype(np.float64)
    assert float(np.abs(got - sharded).max()) <= 2e-5 * scale, float(np.abs(got - sharded).max()) / scale
    assert float(np.abs(got - whole).max()) <= 2e-3 * scale, float(np.abs(got - whole).max()) / scale


# ---------------------------------------------------------------------------------------------------------------------
# SyncBN (optional): batch statistics of MResConv.bn over the rows of all ranks
# ---------------------------------------------------------------------------------------------------------------------
def _cls_net_and_data():
    from meshcnn_b200 import mesh_prepare, synth
    from meshcnn_b200 import networks as N
    torch.manual_seed(0)
    E, B = 270, 3
    net = N.MeshConvNet(N.get_norm_layer('group', 4), 5, [16, 32], 4, E, [200, 150], 10, 1)       # MResConv blocks hold BatchNorm2d
    N.init_weights(net, 'normal', 0.3)
    datas = [mesh_prepare.from_arrays(*synth.make('ico', s, nu=3)) for s in range(B)]
    x = np.stack([d.features for d in datas], 0)
    x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
    y = np.array([0, 3, 1], dtype=np.int64)
    return net, datas, x, y, E, B


def _syncbn_worker(rank, world, port, q, sync):
    import torch.distributed as dist
    from meshcnn_b200 import ddp
    from meshcnn_b200.mesh import Mesh
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    net, datas, x, y, E, B = _cls_net_and_data()
    net = net.to(dev).train()
    ddp.broadcast_parameters(net)
    lo, hi = ddp.shard_range(B, rank, world)
    ddp.enable_sync_batchnorm(sync, local_items=hi - lo, global_items=B)
    red = ddp.FlatGradReducer(net, local_items=hi - lo, global_items=B)
    red.zero()
    out = net(torch.from_numpy(x[lo:hi]).to(dev), [Mesh(data=d) for d in datas[lo:hi]])
    torch.nn.functional.cross_entropy(out, torch.from_numpy(y[lo:hi]).to(dev)).backward()
    red.allreduce()
    torch.cuda.synchronize()
    bn = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)][0]
    q.put((rank, lo, hi, out.detach().cpu().numpy(), red.flat.cpu().numpy(), bn.running_mean.cpu().numpy(), bn.running_var.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sync_batchnorm_makes_the_sharded_step_equal_the_whole_batch_step():
    """With cross-rank BatchNorm statistics the sharded computation IS the whole-batch computation: outputs of every shard
    equal the corresponding rows of the single-process run, the all-reduced gradient equals the whole-batch gradient (to
    the fp32 gradient noise of a network), the running statistics agree.  Without it the per-rank statistics differ
    visibly (uneven shards 2 + 1) -- the switch matters."""
    import torch.multiprocessing as mp
    from meshcnn_b200 import ddp
    from meshcnn_b200.mesh import Mesh
    world = 2
    ctx = mp.get_context('spawn')
    res = {}
    for sync in (True, False):
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_syncbn_worker, args=(r, world, port, q, sync)) for r in range(world)]
        for p in ps:
            p.start()
        got = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
        for p in ps:
            p.join(120)
            assert p.exitcode == 0
        res[sync] = got
    net, datas, x, y, E, B = _cls_net_and_data()
    dev = torch.device('cuda', 0)
    net = net.to(dev).train()
    red = ddp.FlatGradReducer(net)
    red.zero()
    out = net(torch.from_numpy(x).to(dev), [Mesh(data=d) for d in datas])
    torch.nn.functional.cross_entropy(out, torch.from_numpy(y).to(dev)).backward()
    whole_out, whole_grad = out.detach().cpu().numpy(), red.flat.cpu().numpy()
    bn = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)][0]

    def rel(a, b):
        return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / np.abs(b).max())
    sharded_out = np.concatenate([r[3] for r in res[True]], 0)
    assert rel(sharded_out, whole_out) <= 1e-4, rel(sharded_out, whole_out)
    assert rel(res[True][0][4], whole_grad) <= 2e-3, rel(res[True][0][4], whole_grad)
    np.testing.assert_allclose(res[True][0][5], bn.running_mean.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res[True][0][6], bn.running_var.cpu().numpy(), rtol=1e-4, atol=1e-7)
    unsynced_out = np.concatenate([r[3] for r in res[False]], 0)
    assert rel(unsynced_out, whole_out) > 1e-3                       # per-rank statistics: a different function


# ---------------------------------------------------------------------------------------------------------------------
# the graphed multi-GPU step: tail bucket all-reduced from inside backward, head bucket at the end
# ---------------------------------------------------------------------------------------------------------------------
def _graph_worker(rank, world, port, q, mode):
    import torch.distributed as dist
    from meshcnn_b200 import ddp
    from meshcnn_b200.graphed import GraphedTrainStep
    from meshcnn_b200.mesh import Mesh
    from meshcnn_b200.optim import FlatAdam
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    net, datas, x, y, E, B = _net_and_data()
    net = net.to(dev).train()
    ddp.broadcast_parameters(net)
    lo, hi = ddp.shard_range(B, rank, world)
    red = ddp.FlatGradReducer(net, local_items=hi - lo, global_items=B, early_tail=True)
    assert red.split is not None and 0 < red.split < red.flat.numel()
    opt = FlatAdam(red, lr=0.0)                                     # lr 0: the parameters stay put, the gradients are what we compare
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1)
    step = GraphedTrainStep(net, crit, red, opt, [Mesh(data=d, hold_history=True) for d in datas[lo:hi]], x[lo:hi], y[lo:hi], E,
                            collective=mode)
    step([Mesh(data=d, hold_history=True) for d in datas[lo:hi]], x[lo:hi], y[lo:hi])
    torch.cuda.synchronize()
    q.put((rank, red.split, red.flat.cpu().numpy()))
    dist.barrier()
    step.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('mode', ['between', 'in-graph'])
def test_graphed_multi_gpu_step_reduces_the_same_gradient(mode):
    """The captured multi-GPU step in both forms: 'between' (the default: two graphs with the NCCL all-reduce launched
    between them) and 'in-graph' (one graph; with early_tail the tail of the flat buffer -- the late layers -- is reduced
    on a side stream as soon as those gradients are final, the head at the end).  The reduced gradient must equal the
    weighted sum of the shard gradients computed in one process (2e-5), on both ranks, on both sides of the split."""
    import torch.multiprocessing as mp
    from meshcnn_b200 import ddp
    from meshcnn_b200.mesh import Mesh
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_graph_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=900) for _ in range(world)), key=lambda r: r[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res[0][2], res[1][2])
    dev = torch.device('cuda', 0)

    def grad_of(lo, hi):
        net, datas, x, y, E, B = _net_and_data()
        net = net.to(dev).train()
        red = ddp.FlatGradReducer(net)
        red.zero()
        out = net(torch.from_numpy(x[lo:hi]).to(dev), [Mesh(data=d, hold_history=True) for d in datas[lo:hi]])
        torch.nn.CrossEntropyLoss(ignore_index=-1)(out, torch.from_numpy(y[lo:hi]).to(dev)).backward()
        return red.flat.double().cpu().numpy()

    want = (2.0 / 3.0) * grad_of(0, 2) + (1.0 / 3.0) * grad_of(2, 3)
    got, split = res[0][2].astype(np.float64), res[0][1]
    scale = float(np.abs(want).max())
    assert float(np.abs(got[:split] - want[:split]).max()) <= 2e-5 * scale
    assert float(np.abs(got[split:] - want[split:]).max()) <= 2e-5 * scale
