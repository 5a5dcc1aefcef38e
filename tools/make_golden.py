"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on
synthetic meshes.  Run in the build container only:   python tools/make_golden.py

Each fixture stores the inputs and everything the reference produced for them, so the oracle
(oracle/meshcnn_oracle.py) and the CUDA path can be pinned without the reference being present.

Instrumentation of the reference is by wrapping, never editing: ``heappop`` in the mesh_pool module namespace
(pop sequence), ``MeshPool._MeshPool__pool_edge`` (per-pop success), ``MeshUnion.union`` (union log).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

from ref_loader import load_reference, ref_mesh      # noqa: E402
from meshcnn_b200 import mesh_prepare, synth           # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
TMP = tempfile.mkdtemp(prefix='golden_obj_')


def obj_for(kind, seed, **kw):
    vs, faces = synth.make(kind, seed, **kw)
    p = os.path.join(TMP, '%s_%s_%d.obj' % (kind, '_'.join('%s%s' % kv for kv in sorted(kw.items())), seed))
    synth.write_obj(p, vs, faces)
    vs = mesh_prepare.read_obj(p)[0]                 # the values the reference sees (text round trip)
    return p, vs, faces


def ragged(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(l) for l in lists])
    flat = np.array([x for l in lists for x in l], dtype=np.int32)
    return flat, off


def topo(prefix, m, d):
    d[prefix + 'gemm'] = np.array(m.gemm_edges, dtype=np.int32)
    d[prefix + 'sides'] = np.array(m.sides, dtype=np.int8)
    d[prefix + 'edges'] = np.array(m.edges, dtype=np.int32)
    d[prefix + 've_flat'], d[prefix + 've_off'] = ragged(m.ve)
    d[prefix + 'v_mask'] = np.array(m.v_mask, dtype=bool)
    d[prefix + 'ec'] = np.int32(m.edges_count)
    d[prefix + 'vs'] = np.array(m.vs, dtype=np.float64)


class Spy:
    """Wraps the reference's pool internals to record pops / outcomes / unions."""

    def __init__(self, ref):
        self.ref = ref
        self.pops, self.results, self.unions = [], [], []
        mp = ref.mesh_pool
        self._heappop = mp.heappop
        self._pool_edge = mp.MeshPool._MeshPool__pool_edge
        self._union = ref.MeshUnion.union
        spy = self

        def heappop(q):
            item = spy._heappop(q)
            spy.pops.append(int(item[1]))
            spy.results.append(-1)                       # -1 = skipped (masked); overwritten by pool_edge
            return item

        def pool_edge(self_, mesh, edge_id, mask, groups):
            r = spy._pool_edge(self_, mesh, edge_id, mask, groups)
            spy.results[-1] = int(bool(r))
            return r

        def union(self_, s, t):
            spy.unions.append((int(s), int(t)))
            return spy._union(self_, s, t)

        mp.heappop = heappop
        mp.MeshPool._MeshPool__pool_edge = pool_edge
        ref.MeshUnion.union = union

    def reset(self):
        self.pops, self.results, self.unions = [], [], []

    def close(self):
        mp = self.ref.mesh_pool
        mp.heappop = self._heappop
        mp.MeshPool._MeshPool__pool_edge = self._pool_edge
        self.ref.MeshUnion.union = self._union


def rows_to_csr(g):
    """dense clamped [n_new, n_old] -> CSR"""
    g = g.numpy() if torch.is_tensor(g) else g
    ptr = np.zeros(g.shape[0] + 1, dtype=np.int32)
    cols = []
    for r in range(g.shape[0]):
        nz = np.nonzero(g[r])[0]
        cols.append(nz)
        ptr[r + 1] = ptr[r] + len(nz)
    return ptr, (np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32))


def pool_chain_case(ref, name, kind, kw, seed, targets, C, unpool=True, conv_after=True):
    """One mesh through a chain of reference MeshPool levels (+ MeshConv on each pooled level, + MeshUnpool back)."""
    path, vs, faces = obj_for(kind, seed, **kw)
    mesh = ref_mesh(ref, path, hold_history=True)
    d = {'faces': faces.astype(np.int32), 'vs0': vs, 'targets': np.asarray(targets, dtype=np.int32), 'C': np.int32(C)}
    topo('L0_', mesh, d)
    rs = np.random.RandomState(1000 + seed)
    torch.manual_seed(seed)
    E = mesh.edges_count
    fe = torch.from_numpy(rs.randn(1, C, E).astype(np.float32))
    spy = Spy(ref)
    mix = torch.from_numpy((rs.randn(C, C) / np.sqrt(C)).astype(np.float32))
    pooled_feats = []
    for li, T in enumerate(targets):
        spy.reset()
        x = fe.clone().requires_grad_(True)
        ec_in = mesh.edges_count
        norms = torch.sum(x[0, :, :ec_in] * x[0, :, :ec_in], 0).detach().numpy()
        out = ref.MeshPool(T)(x, [mesh])
        dout = torch.from_numpy(rs.randn(*out.shape).astype(np.float32))
        out.backward(dout)
        p = 'P%d_' % li
        d[p + 'x'] = fe.numpy()[0]
        d[p + 'norms'] = norms.astype(np.float32)
        d[p + 'pops'] = np.asarray(spy.pops, dtype=np.int32)
        d[p + 'results'] = np.asarray(spy.results, dtype=np.int8)
        d[p + 'unions'] = np.asarray(spy.unions, dtype=np.int32).reshape(-1, 2)
        d[p + 'out'] = out.detach().numpy()[0]
        d[p + 'dout'] = dout.numpy()[0]
        d[p + 'dx'] = x.grad.numpy()[0]
        g = mesh.history_data['groups'][-1]
        d[p + 'grp_ptr'], d[p + 'grp_col'] = rows_to_csr(g)
        d[p + 'occ'] = mesh.history_data['occurrences'][-1].numpy().astype(np.float32)
        topo(p, mesh, d)
        if conv_after:
            conv = ref.MeshConv(C, C + 2, bias=True)
            xin = out.detach().clone().requires_grad_(True)
            y = conv(xin, [mesh])
            dy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
            y.backward(dy)
            d[p + 'cw'] = conv.conv.weight.detach().numpy()
            d[p + 'cb'] = conv.conv.bias.detach().numpy()
            d[p + 'cy'] = y.detach().numpy()[0]
            d[p + 'cdy'] = dy.numpy()[0]
            d[p + 'cdx'] = xin.grad.numpy()[0]
            d[p + 'cdw'] = conv.conv.weight.grad.numpy()
            d[p + 'cdb'] = conv.conv.bias.grad.numpy()
        pooled_feats.append(out.detach())
        fe = torch.tanh(torch.einsum('oc,bce->boe', mix, out.detach())) + 0.1 * torch.from_numpy(
            rs.randn(*out.shape).astype(np.float32))
    spy.close()
    if unpool:
        sizes = [E] + list(targets)
        f = torch.from_numpy(rs.randn(1, C, targets[-1]).astype(np.float32))
        for ui in range(len(targets) - 1, -1, -1):
            x = f.clone().requires_grad_(True)
            y = ref.MeshUnpool(sizes[ui])(x, [mesh])
            dy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
            y.backward(dy)
            p = 'U%d_' % ui
            d[p + 'x'] = f.numpy()[0]
            d[p + 'y'] = y.detach().numpy()[0]
            d[p + 'dy'] = dy.numpy()[0]
            d[p + 'dx'] = x.grad.numpy()[0]
            d[p + 'gemm_after'] = np.array(mesh.gemm_edges, dtype=np.int32)
            d[p + 'ec_after'] = np.int32(mesh.edges_count)
            f = y.detach()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print(name, 'levels', [int(d['P%d_ec' % i]) for i in range(len(targets))],
          'pops', [len(d['P%d_pops' % i]) for i in range(len(targets))],
          'fails', [int((d['P%d_results' % i] == 0).sum()) for i in range(len(targets))],
          'skips', [int((d['P%d_results' % i] == -1).sum()) for i in range(len(targets))])


def conv_case(ref, name, specs, E, cin, cout, bias):
    """MeshConv forward/backward on a small mixed batch (closed, open-boundary and short meshes padded to E)."""
    rs = np.random.RandomState(7)
    torch.manual_seed(7)
    meshes, d = [], {}
    for i, (kind, kw, seed) in enumerate(specs):
        path, vs, faces = obj_for(kind, seed, **kw)
        m = ref_mesh(ref, path)
        meshes.append(m)
        topo('M%d_' % i, m, d)
        d['M%d_faces' % i] = faces.astype(np.int32)
        d['M%d_vs0' % i] = vs
    B = len(meshes)
    x = torch.from_numpy(rs.randn(B, cin, E).astype(np.float32)).requires_grad_(True)
    conv = ref.MeshConv(cin, cout, bias=bias)
    y = conv(x, meshes)
    dy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    y.backward(dy)
    d.update(B=np.int32(B), E=np.int32(E), x=x.detach().numpy(), w=conv.conv.weight.detach().numpy(),
             y=y.detach().numpy(), dy=dy.numpy(), dx=x.grad.numpy(), dw=conv.conv.weight.grad.numpy())
    if bias:
        d.update(b=conv.conv.bias.detach().numpy(), db=conv.conv.bias.grad.numpy())
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print(name, 'ec', [m.edges_count for m in meshes], 'E', E)


def net_case(ref, name, arch):
    """Tiny end-to-end networks built from the reference's own models/networks.py."""
    import functools
    nets = ref.networks
    torch.manual_seed(11)
    rs = np.random.RandomState(11)
    specs = [('ico', dict(nu=3), 0), ('ico', dict(nu=3), 1), ('torus', dict(m=9, n=10), 2)]
    E = 270
    meshes, d = [], {}
    feats = []
    for i, (kind, kw, seed) in enumerate(specs):
        path, vs, faces = obj_for(kind, seed, **kw)
        m = ref_mesh(ref, path, hold_history=(arch == 'meshunet'))
        meshes.append(m)
        d['M%d_faces' % i] = faces.astype(np.int32)
        d['M%d_vs0' % i] = vs
        f = m.extract_features()
        feats.append(np.pad(f, ((0, 0), (0, E - f.shape[1]))))
    x = np.stack(feats, 0)
    x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
    if arch == 'mconvnet':
        norm = functools.partial(torch.nn.GroupNorm, affine=True, num_groups=4)
        net = nets.MeshConvNet(norm, 5, [8, 16], 4, E, [200, 150], 10, 1)
        labels = torch.tensor([0, 3, 1])
        loss_fn = torch.nn.CrossEntropyLoss()
    else:
        net = nets.MeshEncoderDecoder([E, 200, 150], [5, 8, 16, 32], [32, 16, 8, 3], blocks=1, transfer_data=True)
        labels = torch.from_numpy(rs.randint(0, 3, size=(3, E)))
        loss_fn = torch.nn.CrossEntropyLoss(ignore_index=-1)
    nets.init_weights(net, 'normal', 0.02)
    xt = torch.from_numpy(x).requires_grad_(True)
    out = net(xt, meshes)
    loss = loss_fn(out, labels)
    loss.backward()
    d.update(x=x, labels=labels.numpy(), out=out.detach().numpy(), loss=np.float32(loss.item()), dx=xt.grad.numpy())
    for k, v in net.state_dict().items():
        d['sd/' + k] = v.numpy()
    for k, v in net.named_parameters():
        d['grad/' + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
    print(name, 'loss', float(loss.detach()), 'out', tuple(out.shape), 'ec_end', [m.edges_count for m in meshes])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    conv_case(ref, 'conv_mixed_5_8_bias',
              [('ico', dict(nu=3), 0), ('grid', dict(m=6, n=7), 1), ('torus', dict(m=9, n=10), 2), ('ico', dict(nu=2), 3)],
              E=270, cin=5, cout=8, bias=True)
    conv_case(ref, 'conv_mixed_16_12_nobias',
              [('ico', dict(nu=3), 4), ('grid', dict(m=5, n=9), 5)], E=300, cin=16, cout=12, bias=False)
    pool_chain_case(ref, 'pool_ico750_s0', 'ico', dict(nu=5), 0, [600, 450, 300, 180], C=6)
    pool_chain_case(ref, 'pool_ico750_s1', 'ico', dict(nu=5), 1, [600, 450, 300, 210], C=6)
    pool_chain_case(ref, 'pool_ico750_deep_s3', 'ico', dict(nu=5), 3, [600, 449, 299, 181, 100, 40], C=4)
    pool_chain_case(ref, 'pool_torus750_deep_s4', 'torus', dict(m=10, n=25), 4, [601, 449, 299, 181, 100, 52], C=4)
    pool_chain_case(ref, 'pool_torus2280_s0', 'torus', dict(m=20, n=38), 0, [1800, 1350, 600], C=4)
    pool_chain_case(ref, 'pool_grid_s2', 'grid', dict(m=10, n=12), 2, [330, 280], C=5)
    net_case(ref, 'net_mconvnet_tiny', 'mconvnet')
    net_case(ref, 'net_meshunet_tiny', 'meshunet')


if __name__ == '__main__':
    main()
