"""The host-side surface of the drop-in Mesh / MeshUnion: merge_vertices, remove_edge,
remove_vertex, clean, the history stack (union_groups, remove_group, get_groups, get_occurrences, unroll_gemm, the
old2current / current2old maps) and MeshUnion.rebuild_features / prepare_groups.

Test: the REFERENCE's own, unmodified Python MeshPool / MeshUnpool loops (mesh_pool.py, mesh_unpool.py) are run on CPU over
meshcnn_b200's Mesh and MeshUnion objects, and everything they leave behind is compared bit-exactly with the fixtures the
reference produced on its own Mesh (tests/golden/pool_*.npz).  Needs the reference tree (build container: /root/reference;
elsewhere the copy build() stages into oracle/_ref/)."""
import numpy as np
import pytest
import torch

from oracle import reference_arm
from tests.helpers import assert_topo_equal, golden, prepared, rel_err

pytestmark = pytest.mark.skipif(reference_arm.find_reference() is None, reason='reference tree not present')

CASES = ['pool_ico750_s0', 'pool_ico750_deep_s3', 'pool_grid_s2']


@pytest.fixture()
def ref(monkeypatch):
    from meshcnn_b200.mesh_union import MeshUnion
    r = reference_arm.load()
    monkeypatch.setattr(r.mesh_pool, 'MeshUnion', MeshUnion)          # the reference's pool loop builds OUR MeshUnion
    return r


@pytest.mark.parametrize('name', CASES)
def test_reference_pool_loop_over_our_mesh(ref, name):
    from meshcnn_b200.mesh import Mesh
    d = golden(name)
    mesh = Mesh(data=prepared(d['vs0'], d['faces']), hold_history=True)
    n0 = mesh.edges_count
    targets = d['targets'].tolist()
    for li, T in enumerate(targets):
        p = 'P%d_' % li
        x = torch.from_numpy(d[p + 'x'])[None]                        # [1, C, E]
        out = ref.MeshPool(T)(x, [mesh])
        assert_topo_equal(mesh, d, p)
        np.testing.assert_array_equal(mesh.vs, d[p + 'vs'])
        assert rel_err(out[0].numpy(), d[p + 'out']) < 1e-5
        assert mesh.pool_count == li + 1
        hd = mesh.history_data
        np.testing.assert_array_equal(hd['occurrences'][-1].numpy(), d[p + 'occ'])
        g = hd['groups'][-1].numpy()
        ptr, col = d[p + 'grp_ptr'], d[p + 'grp_col']
        assert g.shape[0] == len(ptr) - 1
        for r in range(g.shape[0]):
            np.testing.assert_array_equal(np.nonzero(g[r])[0], col[ptr[r]:ptr[r + 1]])
        # the maps between original and current edge ids stay inverse of each other on the survivors
        o2c, c2o = hd['old2current'], hd['current2old']
        alive = np.nonzero(o2c != -1)[0]
        assert len(alive) == mesh.edges_count and len(o2c) == n0
        np.testing.assert_array_equal(c2o[o2c[alive]], alive)
    sizes = [int(d['L0_ec'])] + targets
    for ui in range(len(targets) - 1, -1, -1):
        p = 'U%d_' % ui
        if p + 'x' not in d.files:
            break
        x = torch.from_numpy(d[p + 'x'])[None]
        y = ref.MeshUnpool(sizes[ui])(x, [mesh])
        assert rel_err(y[0].numpy(), d[p + 'y']) < 1e-5
        np.testing.assert_array_equal(mesh.gemm_edges, d[p + 'gemm_after'])
        assert mesh.edges_count == int(d[p + 'ec_after'])


def test_meshunion_matches_reference_dense():
    """Sparse-row MeshUnion == the reference's dense n x n matrix under a random union sequence (multiplicities, F6)."""
    from meshcnn_b200.mesh_union import MeshUnion
    r = reference_arm.load()
    rs = np.random.RandomState(0)
    n = 40
    a, b = MeshUnion(n), r.MeshUnion(n)
    for _ in range(120):
        s, t = rs.randint(0, n, 2)
        a.union(int(s), int(t))
        b.union(int(s), int(t))
    np.testing.assert_array_equal(a.get_occurrences().numpy(), b.get_occurrences().numpy())
    mask = rs.rand(n) > 0.4
    f = torch.from_numpy(rs.randn(6, n + 5).astype(np.float32))
    np.testing.assert_array_equal(a.get_group(3).numpy(), b.get_group(3).numpy())
    fa = a.rebuild_features(f, mask, 30)
    fb = b.rebuild_features(f, mask, 30)
    assert fa.shape == fb.shape
    np.testing.assert_allclose(fa.numpy(), fb.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(a.groups.numpy(), b.groups.numpy())     # post-state of prepare_groups: [E, n_surviving]
    c, e = MeshUnion(n), r.MeshUnion(n)
    for s, t in [(1, 2), (1, 2), (2, 3)]:
        c.union(s, t)
        e.union(s, t)
    tm = torch.from_numpy(mask)
    np.testing.assert_array_equal(c.get_groups(tm).numpy(), e.get_groups(tm).numpy())
    np.testing.assert_array_equal(c.groups.numpy(), e.groups.numpy())
