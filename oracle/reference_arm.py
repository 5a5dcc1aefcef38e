"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product (meshcnn_b200/).

Drives the UNMODIFIED reference (ranahanocka/MeshCNN) on the host CPU for bench.py's `--impl reference` arm and its
`cpu_baseline` object, and for parity tests that want the reference's own numbers.

Where the reference comes from: `/root/reference` in the build container; on the GPU box (no such path) the staged copy
`oracle/_ref/reference/` that `__graft_entry__.build()` makes from it (`oracle/_ref/` is git-ignored, travels with gpurun).
Nothing is edited: the two incompatibilities with the container's numpy (`np.bool` alias, ragged `ve` inside
`np.savez_compressed`) are shimmed from outside, and timers are installed by wrapping `MeshPool.forward`.

What one step is (models/mesh_classifier.py:56-68 of the reference): `optimizer.zero_grad(); out = net(edge_features, mesh);
loss = criterion(out, labels); loss.backward(); optimizer.step()` with `net = networks.define_classifier(...)` (networks.py:96-111)
and `torch.optim.Adam` (mesh_classifier.py:38), on fresh single-use `Mesh` objects made outside the timed region.
"""
import copy
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.environ.get('MESHCNN_REFERENCE') or '/root/reference', os.path.join(HERE, '_ref', 'reference'))
STAGED = os.path.join(HERE, '_ref', 'reference')
STAGE_ITEMS = ('models', 'util', 'options', 'data', 'train.py', 'test.py', 'LICENSE')


def find_reference():
    for p in CANDIDATES:
        if p and os.path.isfile(os.path.join(p, 'models', 'layers', 'mesh_pool.py')):
            return p
    return None


def stage(src='/root/reference'):
    """Copy the reference's Python packages to oracle/_ref/reference (build container only; a no-op elsewhere)."""
    import shutil
    if not os.path.isdir(src):
        return False
    os.makedirs(STAGED, exist_ok=True)
    for it in STAGE_ITEMS:
        s, d = os.path.join(src, it), os.path.join(STAGED, it)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', 'cache'))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    return True


_REF = None


def load():
    """The reference's modules under their own dotted names (`models.layers.*`, `models.networks`)."""
    global _REF
    if _REF is not None:
        return _REF
    root = find_reference()
    if root is None:
        raise RuntimeError('reference tree not found (looked in %s); run __graft_entry__.build() in the build container '
                           'to stage it into oracle/_ref/' % (CANDIDATES,))
    if root not in sys.path:
        sys.path.insert(0, root)
    import numpy as np
    if not hasattr(np, 'bool'):                      # mesh_pool.py:47 uses the alias numpy removed
        np.bool = bool
    import models.layers.mesh as mesh_mod
    import models.layers.mesh_conv as mc
    import models.layers.mesh_pool as mpool
    import models.layers.mesh_prepare as mp
    import models.layers.mesh_union as munion
    import models.layers.mesh_unpool as munpool
    import models.networks as networks

    def fill_mesh_nocache(mesh2fill, file, opt):     # fill_mesh minus the npz cache; same fields, same source
        d = mp.from_scratch(file, opt)
        for k in ('vs', 'edges', 'gemm_edges', 've', 'v_mask', 'edge_lengths', 'edge_areas', 'features', 'sides'):
            setattr(mesh2fill, k, getattr(d, k))
        mesh2fill.edges_count = int(d.edges_count)
        mesh2fill.filename = str(d.filename)

    mesh_mod.fill_mesh = fill_mesh_nocache
    _REF = types.SimpleNamespace(root=root, mesh_prepare=mp, mesh=mesh_mod, mesh_conv=mc, mesh_pool=mpool, mesh_unpool=munpool,
                                 mesh_union=munion, networks=networks, Mesh=mesh_mod.Mesh, MeshConv=mc.MeshConv,
                                 MeshPool=mpool.MeshPool, MeshUnpool=munpool.MeshUnpool, MeshUnion=munion.MeshUnion)
    return _REF


def reference_meshes(cfg, B, first_seed, hold_history):
    """B reference `Mesh` objects of the workload's synthetic meshes (written as .obj, read by the reference's own
    mesh_prepare.from_scratch) + the standardised, padded feature tensor [B, 5, E] the data loader would hand over."""
    import numpy as np
    from meshcnn_b200 import synth
    ref = load()
    tmp = tempfile.mkdtemp(prefix='mcnn_ref_obj_')
    opt = types.SimpleNamespace(num_aug=1)
    meshes = []
    for i in range(B):
        s = first_seed + i
        kind, kw = cfg['mesh'] if s % 2 == 0 else cfg['mesh2']
        vs, faces = synth.make(kind, s, **kw)
        p = os.path.join(tmp, '%s_%d.obj' % (kind, s))
        synth.write_obj(p, vs, faces)
        meshes.append(ref.Mesh(file=p, opt=opt, hold_history=hold_history))
        os.remove(p)
    os.rmdir(tmp)
    E = cfg['E']
    x = np.zeros((B, 5, E), dtype=np.float64)
    for i, m in enumerate(meshes):
        x[i, :, :m.edges_count] = m.features
    x = ((x - x.mean(axis=(0, 2), keepdims=True)) / x.std(axis=(0, 2), keepdims=True)).astype(np.float32)
    return meshes, x


def build_net(cfg):
    """networks.define_classifier (networks.py:96-111) with the flags of the reference's scripts/*/train.sh."""
    import torch
    ref = load()
    opt = types.SimpleNamespace(norm='group', num_groups=cfg.get('groups', 16), pool_res=list(cfg['pool_res']), fc_n=cfg.get('fc_n', 100),
                                resblocks=cfg['resblocks'], dataset_mode='classification' if cfg['arch'] == 'mconvnet' else 'segmentation')
    torch.manual_seed(0)
    net = ref.networks.define_classifier(5, list(cfg['ncf']), cfg['E'], cfg['nclasses'], opt, [], cfg['arch'], 'normal', 0.02)
    return net, ref.networks.define_loss(opt)


class _PoolTimer:
    """Accumulates the wall time of every reference MeshPool.forward call (wrapped, not edited)."""

    def __init__(self, ref):
        self.cls = ref.MeshPool
        self.orig = ref.MeshPool.forward
        self.t = 0.0
        timer = self

        def forward(self_, fe, meshes):
            t0 = time.perf_counter()
            try:
                return timer.orig(self_, fe, meshes)
            finally:
                timer.t += time.perf_counter() - t0
        self.cls.forward = forward

    def close(self):
        self.cls.forward = self.orig


def run(cfg, B, steps, warmup, first_seed=0, budget_s=None, threads=None):
    """Times `steps` training steps (after `warmup`) of the unmodified reference on B meshes of the workload `cfg` (a
    bench.py CONFIGS entry).  With `budget_s` the number of timed steps is chosen from the first timed step so that the
    sample is about that many seconds of CPU work (at least `steps`, at most 40).
    Returns a dict: meshes_per_s, cores, s_per_step, fwd_s, bwd_s, opt_s, pool_s (per step means), steps, B."""
    import numpy as np
    import torch
    ref = load()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    hold = cfg['arch'] == 'meshunet'
    pristine, x = reference_meshes(cfg, B, first_seed, hold)
    rs = np.random.RandomState(first_seed)
    if cfg['arch'] == 'mconvnet':
        labels = rs.randint(0, cfg['nclasses'], size=(B,)).astype(np.int64)
    else:
        labels = rs.randint(0, cfg['nclasses'], size=(B, cfg['E'])).astype(np.int64)
        for i, m in enumerate(pristine):
            labels[i, m.edges_count:] = -1
    net, crit = build_net(cfg)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=cfg['lr'], betas=(0.9, 0.999))          # mesh_classifier.py:38
    xt, lt = torch.from_numpy(x), torch.from_numpy(labels)
    pt = _PoolTimer(ref)
    acc = {'fwd': 0.0, 'bwd': 0.0, 'opt': 0.0, 'pool': 0.0, 'n': 0}
    try:
        it = -1
        while True:
            it += 1
            if it >= warmup + steps:
                break
            meshes = np.array([copy.deepcopy(m) for m in pristine], dtype=object)       # fresh (F11), object array (F12); untimed
            xin = xt.clone().requires_grad_(True)                                      # mesh_classifier.py:49
            pt.t = 0.0
            t0 = time.perf_counter()
            opt.zero_grad()
            out = net(xin, meshes)
            loss = crit(out, lt)
            t1 = time.perf_counter()
            loss.backward()
            t2 = time.perf_counter()
            opt.step()
            t3 = time.perf_counter()
            if it >= warmup:
                acc['fwd'] += t1 - t0
                acc['bwd'] += t2 - t1
                acc['opt'] += t3 - t2
                acc['pool'] += pt.t
                acc['n'] += 1
                if budget_s is not None and acc['n'] == 1:
                    steps = max(steps, min(40, int(round(budget_s / max(t3 - t0, 1e-3)))))
    finally:
        pt.close()
    n = acc['n']
    total = acc['fwd'] + acc['bwd'] + acc['opt']
    return {'meshes_per_s': B * n / total, 'cores': threads, 's_per_step': total / n, 'fwd_s': acc['fwd'] / n, 'bwd_s': acc['bwd'] / n,
            'opt_s': acc['opt'] / n, 'pool_s': acc['pool'] / n, 'steps': n, 'B': B, 'loss': float(loss.detach()),
            'root': ref.root}
