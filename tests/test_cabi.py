"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/meshcnn_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from meshcnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header='meshcnn_b200.h'):
    src = open(os.path.join(ROOT, 'include', header)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mcnn_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_all_bound_and_exported():
    names = _declared()
    assert len(names) >= 14
    assert sorted(_lib.SIGNATURES) == names, 'ctypes table and header disagree'
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('library not built yet (run __graft_entry__.build())')
    so = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), n
    assert _lib.lib().mcnn_version() >= 100


def test_debug_switches_live_in_their_own_header():
    """The process-global diagnostics knobs are not part of the production ABI: separate header, separate ctypes table."""
    prod, dbg = _declared(), _declared('meshcnn_b200_debug.h')
    assert not [n for n in prod if n.startswith('mcnn_debug_')]
    assert dbg and all(n.startswith('mcnn_debug_') for n in dbg)
    assert sorted(_lib.DEBUG_SIGNATURES) == dbg
    if os.path.exists(_lib.LIB_PATH):
        so = ctypes.CDLL(_lib.LIB_PATH)
        for n in dbg:
            assert hasattr(so, n), n


def test_pool_args_struct_matches_header_size():
    # 8 int32 + 4 ptr + (ptr, 2 int32) + 4 ptr + 3 ptr + 6 ptr + 6 ptr on LP64
    assert ctypes.sizeof(_lib.PoolArgs) == 8 * 4 + 4 * 8 + 8 + 8 + 4 * 8 + 3 * 8 + 6 * 8 + 7 * 8 + 3 * 8


def test_smem_budget_queries():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('library not built yet')
    L = _lib.lib()
    lim = L.mcnn_pool_smem_limit()
    assert 200 * 1024 <= lim <= 227 * 1024                              # 227 KB per CTA minus the kernel's static shared memory
    assert L.mcnn_pool_smem_bytes(750, 252, 1500, 0) < 64 * 1024
    assert L.mcnn_pool_smem_bytes(2280, 760, 4560, 1) < lim            # human-seg / COSEG resolution fits on chip
    assert L.mcnn_pool_smem_bytes(170000, 57000, 340000, 0) > lim      # config 5 does not


def test_layers_refuse_cpu_tensors():
    """No CPU fallback: a CPU tensor must fail loudly, not silently compute somewhere else."""
    import numpy as np
    import torch
    from meshcnn_b200 import synth, mesh_prepare
    from meshcnn_b200.mesh import Mesh
    from meshcnn_b200.mesh_conv import MeshConv
    vs, faces = synth.make('ico', 0, nu=2)
    m = Mesh(data=mesh_prepare.from_arrays(vs, faces))
    conv = MeshConv(5, 4)
    with pytest.raises(Exception):
        conv(torch.zeros(1, 5, m.edges_count), [m])
