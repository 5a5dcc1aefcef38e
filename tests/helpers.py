"""Shared helpers for the parity tests: fixture loading and oracle mesh construction."""
import glob
import os

import numpy as np

from meshcnn_b200 import mesh_prepare

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_names(prefix):
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def unragged(flat, off):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def prepared(vs, faces, name='golden'):
    return mesh_prepare.from_arrays(vs, faces, filename=name)


def assert_topo_equal(mesh, d, prefix, check_ve=True):
    """mesh: object with gemm_edges/sides/edges/ve/v_mask/edges_count; d: fixture; bit-exact comparison."""
    ec = int(d[prefix + 'ec'])
    assert int(mesh.edges_count) == ec, (mesh.edges_count, ec)
    np.testing.assert_array_equal(np.asarray(mesh.gemm_edges)[:len(d[prefix + 'gemm'])], d[prefix + 'gemm'])
    np.testing.assert_array_equal(np.asarray(mesh.sides)[:len(d[prefix + 'sides'])], d[prefix + 'sides'])
    np.testing.assert_array_equal(np.asarray(mesh.edges)[:len(d[prefix + 'edges'])], d[prefix + 'edges'])
    np.testing.assert_array_equal(np.asarray(mesh.v_mask), d[prefix + 'v_mask'])
    if check_ve:
        assert [list(map(int, l)) for l in mesh.ve] == unragged(d[prefix + 've_flat'], d[prefix + 've_off'])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def close(a, b, rtol, atol=0.0):
    """max|a-b| <= rtol * max|b| + atol  (a norm-wise relative check; atol absorbs gradients that are
    mathematically zero, e.g. a conv bias feeding an InstanceNorm)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) <= rtol * float(np.abs(b).max()) + atol
