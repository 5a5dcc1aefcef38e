"""GraphedTrainStep (whole step replayed as one CUDA graph) against the same steps launched kernel by kernel: losses,
parameters and the pooled topology of the last batch must agree.  Also MeshBatch.load (device buffers re-used for new
meshes)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(arch):
    from meshcnn_b200 import mesh_prepare, synth
    from meshcnn_b200 import networks as N
    torch.manual_seed(0)
    B, E = 4, 270
    if arch == 'mconvnet':
        net = N.MeshConvNet(N.get_norm_layer('group', 4), 5, [16, 32], 4, E, [200, 150], 10, 1)
    else:
        net = N.MeshEncoderDecoder([E, 200, 150], [5, 16, 32, 64], [64, 32, 16, 4], blocks=1, transfer_data=True)
    N.init_weights(net, 'normal', 0.02)
    batches = []
    for k in range(3):
        datas = [mesh_prepare.from_arrays(*synth.make('ico', 10 * k + s, nu=3)) for s in range(B)]
        x = np.stack([d.features for d in datas], 0)
        x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
        rs = np.random.RandomState(k)
        y = rs.randint(0, 4, size=(B,) if arch == 'mconvnet' else (B, E)).astype(np.int64)
        batches.append((datas, x, y))
    return net, batches, E


@pytest.mark.parametrize('arch', ['mconvnet', 'meshunet'])
def test_graph_replay_matches_eager(arch):
    """Three training steps, each started from identical parameters and Adam moments (copied from the eager run, so that
    the rounding noise of the scatter atomics cannot compound through Adam's normalisation or flip a near-tie in the
    collapse order): loss, flat gradient, updated parameters and pooled topology must agree."""
    from meshcnn_b200 import ddp
    from meshcnn_b200.graphed import GraphedTrainStep
    from meshcnn_b200.mesh import Mesh
    from meshcnn_b200.optim import FlatAdam
    from tests.helpers import close
    dev = torch.device('cuda:0')
    net0, batches, E = _setup(arch)
    hold = arch == 'meshunet'
    crit = torch.nn.CrossEntropyLoss()
    ne = copy.deepcopy(net0).to(dev).train()
    re_ = ddp.FlatGradReducer(ne)
    oe = FlatAdam(re_, lr=1e-3)
    ng = copy.deepcopy(net0).to(dev).train()
    rg = ddp.FlatGradReducer(ng)
    og = FlatAdam(rg, lr=1e-3)
    d0, x0, y0 = batches[0]
    step = GraphedTrainStep(ng, crit, rg, og, [Mesh(data=d, hold_history=hold) for d in d0], x0, y0, E, source='host')
    for k, (datas, x, y) in enumerate(batches):
        # graph side starts from the eager side's state (capture + warm-up ran real optimiser steps)
        ng.load_state_dict(ne.state_dict())
        og.exp_avg.copy_(oe.exp_avg); og.exp_avg_sq.copy_(oe.exp_avg_sq)
        og.state[5:6].fill_(float(k))
        me = [Mesh(data=d, hold_history=hold) for d in datas]
        xin = torch.from_numpy(x).to(dev).requires_grad_(True)
        re_.zero()
        le = crit(ne(xin, me), torch.from_numpy(y).to(dev))
        le.backward()
        ge = re_.flat.clone()
        oe.step()
        mg = [Mesh(data=d, hold_history=hold) for d in datas]
        lg = step(mg, x, y)
        assert abs(float(lg.item()) - float(le.item())) <= 1e-5 * abs(float(le.item())) + 1e-6
        assert close(rg.flat.cpu().numpy(), ge.cpu().numpy(), 1e-4, 1e-7)
        assert close(step.x_dev.grad.cpu().numpy(), xin.grad.cpu().numpy(), 1e-4, 1e-7)
        # Adam divides by sqrt(v): an element whose gradient is pure rounding noise (e.g. a conv bias in front of an
        # InstanceNorm: mathematically zero) moves by an arbitrary fraction of lr, so only elements with a significant
        # gradient are compared here; the kernel itself is checked exactly in the test below
        gnp = ge.cpu().numpy()
        sig = np.abs(gnp) > 1e-2 * np.abs(gnp).max()
        assert sig.sum() > 0
        assert np.abs(og.flat_p.cpu().numpy()[sig] - oe.flat_p.cpu().numpy()[sig]).max() <= 2e-2 * 1e-3
        for a, b in zip(me, mg):                      # host mirrors refresh lazily from the replayed device state
            assert int(a.edges_count) == int(b.edges_count)
            np.testing.assert_array_equal(a.gemm_edges, b.gemm_edges)


def test_back_to_back_steps_do_not_race_on_the_staging_buffers():
    """Steps issued without reading the loss in between: every replay must still see ITS batch (stage() waits for the
    previous replay's host->device copies before it rewrites the pinned buffers).  Building the step must not train:
    the warm-up's parameter / Adam / BatchNorm updates are undone."""
    from meshcnn_b200 import ddp
    from meshcnn_b200.graphed import GraphedTrainStep
    from meshcnn_b200.mesh import Mesh
    from meshcnn_b200.optim import FlatAdam
    dev = torch.device('cuda:0')
    net0, batches, E = _setup('mconvnet')
    crit = torch.nn.CrossEntropyLoss()
    losses = []
    for mode in ('back_to_back', 'synchronous'):
        net = copy.deepcopy(net0).to(dev).train()
        red = ddp.FlatGradReducer(net)
        opt = FlatAdam(red, lr=2e-4)
        before = opt.flat_p.clone()
        bn_before = [b.clone() for b in net.buffers()]
        datas, x, y = batches[0]
        step = GraphedTrainStep(net, crit, red, opt, [Mesh(data=d) for d in datas], x, y, E)
        torch.testing.assert_close(opt.flat_p, before, rtol=0, atol=0)           # construction did not train
        assert float(opt.state[5]) == 0.0 and opt.step_count == 0 and float(opt.exp_avg.abs().max()) == 0.0
        for b, v in zip(net.buffers(), bn_before):
            torch.testing.assert_close(b, v, rtol=0, atol=0)
        got = []
        for rep in range(4):
            for datas, x, y in batches:
                got.append(step([Mesh(data=d) for d in datas], x, y).clone())
                if mode == 'synchronous':
                    torch.cuda.synchronize()
        torch.cuda.synchronize()
        losses.append(torch.stack(got).cpu())
    torch.testing.assert_close(losses[0], losses[1], rtol=2e-3, atol=1e-5)
    assert float((losses[0][:3] - losses[0][3:6]).abs().max()) > 0              # the three batches do differ / training moves


def test_adam_state_kernel_matches_host_driven_kernel():
    """mcnn_adam_flat_state (hyper-parameters + step counter in device memory) == mcnn_adam_flat, step by step."""
    from meshcnn_b200 import ddp
    from meshcnn_b200.optim import FlatAdam
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    lin = [torch.nn.Linear(37, 19).to(dev) for _ in range(2)]
    lin[1].load_state_dict(lin[0].state_dict())
    reds = [ddp.FlatGradReducer(m) for m in lin]
    opts = [FlatAdam(r, lr=3e-3, betas=(0.9, 0.999), weight_decay=0.01) for r in reds]
    for it in range(5):
        g = torch.randn_like(reds[0].flat)
        for r in reds:
            r.flat.copy_(g)
        opts[0].step()
        opts[1].step_state()
        if it == 2:
            for o in opts:
                o.param_groups[0]['lr'] = 1e-3
            opts[1].sync_lr()
        torch.testing.assert_close(opts[1].flat_p, opts[0].flat_p, rtol=1e-6, atol=1e-8)


def test_meshbatch_load_reuses_buffers():
    from meshcnn_b200 import mesh_prepare, synth
    from meshcnn_b200.mesh import Mesh, MeshBatch
    from meshcnn_b200.mesh_pool import MeshPool
    dev = torch.device('cuda:0')
    E = 270
    da = [mesh_prepare.from_arrays(*synth.make('ico', s, nu=3)) for s in range(2)]
    db = [mesh_prepare.from_arrays(*synth.make('ico', 7 + s, nu=3)) for s in range(2)]
    torch.manual_seed(1)
    f = torch.randn(2, E, 8, device=dev)
    ref = [Mesh(data=d) for d in db]
    out_ref = MeshPool(200).forward_em(f, MeshBatch(ref, E, dev))
    ma = [Mesh(data=d) for d in da]
    bt = MeshBatch(ma, E, dev)
    MeshPool(200).forward_em(f, bt)
    ptr = bt.blob.data_ptr()
    mb = [Mesh(data=d) for d in db]
    bt.load(mb)
    assert bt.blob.data_ptr() == ptr and bt.cur == 0
    out = MeshPool(200).forward_em(f, bt)
    bt.check()
    torch.testing.assert_close(out, out_ref, rtol=0, atol=0)
    for a, b in zip(ref, mb):
        np.testing.assert_array_equal(a.gemm_edges, b.gemm_edges)
        assert a.ve == b.ve
