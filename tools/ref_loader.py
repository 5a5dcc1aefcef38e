"""Import the unmodified reference (ranahanocka/MeshCNN) from /root/reference for golden-vector generation.

Only usable in the build container (the GPU box has no /root/reference).  Nothing under tests/ with the
``gpu`` marker, bench.py or __graft_entry__.smoke() imports this file.

The one shim needed: numpy >= 1.24 refuses to build a ragged array from ``ve`` inside
``np.savez_compressed``; we bypass the npz cache by replacing ``fill_mesh`` with a version that calls the
reference's own ``from_scratch`` and copies the same fields.  No reference file is modified or copied.
"""
import os
import sys
import types

REF = os.environ.get('MESHCNN_REFERENCE', '/root/reference')


def load_reference():
    """Returns a namespace with the reference's layer modules (imported under their own dotted names)."""
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree not found at %s' % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import numpy as np
    if not hasattr(np, 'bool'):          # mesh_pool.py:47 uses the removed alias on numpy 1.24-1.26 / 2.x
        np.bool = bool
    import models.layers.mesh_prepare as mp
    import models.layers.mesh as mesh_mod
    import models.layers.mesh_conv as mc
    import models.layers.mesh_pool as mpool
    import models.layers.mesh_unpool as munpool
    import models.layers.mesh_union as munion
    import models.networks as networks

    def fill_mesh_nocache(mesh2fill, file, opt):
        d = mp.from_scratch(file, opt)
        for k in ('vs', 'edges', 'gemm_edges', 've', 'v_mask', 'edge_lengths', 'edge_areas', 'features', 'sides'):
            setattr(mesh2fill, k, getattr(d, k))
        mesh2fill.edges_count = int(d.edges_count)
        mesh2fill.filename = str(d.filename)

    mesh_mod.fill_mesh = fill_mesh_nocache
    return types.SimpleNamespace(mesh_prepare=mp, mesh=mesh_mod, mesh_conv=mc, mesh_pool=mpool,
                                 mesh_unpool=munpool, mesh_union=munion, networks=networks,
                                 Mesh=mesh_mod.Mesh, MeshConv=mc.MeshConv, MeshPool=mpool.MeshPool,
                                 MeshUnpool=munpool.MeshUnpool, MeshUnion=munion.MeshUnion)


def ref_mesh(ref, obj_path, hold_history=False):
    opt = types.SimpleNamespace(num_aug=1)
    return ref.Mesh(file=obj_path, opt=opt, hold_history=hold_history)
