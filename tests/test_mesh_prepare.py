"""meshcnn_b200.mesh_prepare against fixtures made by the unmodified reference's mesh_prepare.from_scratch
(tools/make_golden_prepare.py): topology, the five edge features, edge areas / lengths -- and the data augmentation
(scale_verts / flip_edges / slide_verts), which draws from the global np.random in the reference's order, so that an equal
seed gives a bit-identical augmented mesh.  Bit-exact comparisons throughout (float64 arrays compared with array_equal)."""
import os
import types

import numpy as np
import pytest

from meshcnn_b200 import mesh_prepare as MP
from meshcnn_b200 import synth
from tests.helpers import golden, golden_names, unragged

NAMES = golden_names('prepare_')


def _opt(d):
    kw = {k: v for k, v in zip(d['opt_keys'].tolist(), d['opt_vals'].tolist())}
    kw['num_aug'] = int(kw['num_aug'])
    return types.SimpleNamespace(**kw)


@pytest.mark.parametrize('name', NAMES)
def test_prepare_matches_reference(name):
    d = golden(name)
    np.random.seed(int(d['seed']))
    m = MP.from_arrays(d['vs_in'], d['faces_in'], opt=_opt(d))
    assert np.random.randint(0, 1 << 30) == int(d['next_draw'])          # same number of draws, in the same order
    assert m.edges_count == int(d['ec'])
    np.testing.assert_array_equal(m.vs, d['vs'])
    np.testing.assert_array_equal(m.edges, d['edges'])
    np.testing.assert_array_equal(m.gemm_edges, d['gemm'])
    np.testing.assert_array_equal(m.sides, d['sides'])
    assert [list(map(int, l)) for l in m.ve] == unragged(d['ve_flat'], d['ve_off'])
    np.testing.assert_array_equal(m.edge_areas, d['edge_areas'])
    assert m.edge_areas.dtype == d['edge_areas'].dtype
    np.testing.assert_array_equal(m.edge_lengths, d['edge_lengths'])
    np.testing.assert_array_equal(m.features, d['features'])


def test_fixtures_cover_flips_and_slides():
    moved = [float(golden(n)['shifted']) for n in NAMES if n.endswith('_slide')]
    assert max(moved) >= 0.4
    d0, d1 = golden('prepare_ico3_plain'), golden('prepare_ico3_flip')
    assert (d0['edges'] != d1['edges']).any()


def test_mesh_honours_opt_and_cache(tmp_path):
    """Mesh(file, opt) -> fill_mesh semantics (mesh.py:16, mesh_prepare.py:6-37): augmentation per opt, one npz per
    augmentation slot under <dir>/cache/, reloaded on a hit."""
    from meshcnn_b200.mesh import Mesh
    vs, faces = synth.icosphere(3)
    p = str(tmp_path / 'sphere.obj')
    synth.write_obj(p, vs * 50, faces)
    opt = types.SimpleNamespace(num_aug=3, scale_verts=True, flip_edges=0.2, slide_verts=0.2)
    np.random.seed(5)
    a = Mesh(file=p, opt=opt)
    files = sorted(os.listdir(tmp_path / 'cache'))
    assert len(files) == 1 and files[0].startswith('sphere_00') and files[0].endswith('.npz')
    plain = Mesh(file=p, opt=None)
    assert not np.array_equal(a.vs, plain.vs)                            # the augmentation was applied
    z = np.load(str(tmp_path / 'cache' / files[0]), allow_pickle=True)
    assert sorted(z.files) == sorted(MP.CACHE_KEYS)                       # the reference's key set
    np.random.seed(5)                                                     # same slot again -> served from the cache
    b = Mesh(file=p, opt=opt)
    assert sorted(os.listdir(tmp_path / 'cache')) == files
    for k in ('vs', 'edges', 'gemm_edges', 'sides', 'features', 'edge_areas', 'edge_lengths'):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    assert a.ve == b.ve and a.edges_count == b.edges_count and a.filename == b.filename == 'sphere.obj'
    for _ in range(12):                                                   # other slots get their own files (at most num_aug)
        Mesh(file=p, opt=opt)
    assert 2 <= len(os.listdir(tmp_path / 'cache')) <= 3


def test_native_topology_builder_matches_the_numpy_formulation():
    """csrc/mesh_topology.cu (mcnn_host_drop_non_manifold / mcnn_host_build_gemm) against the NumPy / Python formulation of the
    same steps, on clean meshes, on input with zero-area and directed-edge-repeating faces, and on the >2-faces-per-edge error."""
    import numpy as np
    from meshcnn_b200 import mesh_prepare, synth
    cases = [synth.make('ico', 1, nu=5), synth.make('torus', 2, m=10, n=25), synth.make('torus', 5, m=7, n=9)]
    vs, faces = synth.make('ico', 3, nu=3)
    faces = np.concatenate([faces, faces[:5], faces[7:9][:, ::-1], np.array([[0, 0, 1], [2, 3, 3]])])   # repeats, flips, zero area
    cases.append((vs, faces))
    vs, faces = synth.make('torus', 4, m=6, n=8)
    cases.append((vs, faces[:-17]))                                                                    # an open mesh (boundaries)
    for vs, faces in cases:
        vs = np.asarray(vs, dtype=np.float64)
        faces = np.asarray(faces, dtype=np.int64)
        f1, a1 = mesh_prepare.drop_non_manifold(vs, faces)
        f0, a0 = mesh_prepare.drop_non_manifold_py(vs, faces)
        assert np.array_equal(f1, f0) and np.array_equal(a1, a0)
        t1 = mesh_prepare.build_topology(vs, f1, a1)
        t0 = mesh_prepare.build_topology_py(vs, f0, a0)
        for x, y in zip(t1[:3], t0[:3]):
            assert x.dtype == y.dtype and np.array_equal(x, y)
        assert t1[3] == t0[3]
        assert t1[4].dtype == t0[4].dtype and np.array_equal(t1[4], t0[4])
        # the features: everything up to the arccos natively (mcnn_host_edge_features), bit-identical to the NumPy formulation
        f1, l1 = mesh_prepare.extract_features_native(vs, t1[0], t1[1])
        f0, l0 = mesh_prepare.extract_features_py(vs, t0[0], t0[1])
        assert f1.dtype == f0.dtype and f1.shape == f0.shape and np.array_equal(f1, f0, equal_nan=True)
        assert np.array_equal(l1, l0)
    vs = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]])
    fan = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4]])              # the undirected edge (0, 1) in three faces
    for fn in (mesh_prepare.build_topology, mesh_prepare.build_topology_py):
        with pytest.raises(ValueError):
            fn(vs, fan)


def test_native_features_match_numpy_on_a_collapsed_edge():
    """An edge whose endpoints coincide: base length 0 and projection 0, so the height / base ratio is 0 / 0 -- NaN, not a
    divide error -- in the NumPy formulation; the native path must produce the same NaNs in the same places."""
    import warnings
    import numpy as np
    from meshcnn_b200 import mesh_prepare, synth
    vs, faces = synth.make('ico', 1, nu=3)
    vs = np.asarray(vs, dtype=np.float64)
    f, a = mesh_prepare.drop_non_manifold(vs, np.asarray(faces, dtype=np.int64))
    edges, gemm = mesh_prepare.build_topology(vs, f, a)[:2]
    vs = vs.copy()
    vs[edges[5, 1]] = vs[edges[5, 0]]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        f1, l1 = mesh_prepare.extract_features_native(vs, edges, gemm)
        f0, l0 = mesh_prepare.extract_features_py(vs, edges, gemm)
    assert np.isnan(f0).any()
    assert np.array_equal(f1, f0, equal_nan=True) and np.array_equal(l1, l0)
