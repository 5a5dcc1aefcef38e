"""meshcnn_b200: B200-native (sm_100a) implementation of MeshCNN's per-edge hot path -- MeshConv, MeshPool /
MeshUnion, MeshUnpool -- behind the reference's own layer API.  See DESIGN.md."""
from . import _lib  # noqa: F401

__all__ = ['Mesh', 'MeshBatch', 'MeshConv', 'MeshPool', 'MeshUnpool', 'MeshUnion', 'set_precision']


def __getattr__(name):
    if name in ('Mesh', 'MeshBatch'):
        from . import mesh
        return getattr(mesh, name)
    if name == 'MeshConv':
        from .mesh_conv import MeshConv
        return MeshConv
    if name == 'MeshPool':
        from .mesh_pool import MeshPool
        return MeshPool
    if name == 'MeshUnpool':
        from .mesh_unpool import MeshUnpool
        return MeshUnpool
    if name == 'set_precision':
        from .functional import set_precision
        return set_precision
    if name == 'MeshUnion':
        from .mesh_union import MeshUnion
        return MeshUnion
    raise AttributeError(name)
