"""Drop-in for models/layers/mesh_unpool.py: the dense [B, E_lo, E_hi] unroll matrix becomes a CSC gather."""
import torch.nn as nn

from .functional import UnpoolFn
from .layout import as_bce, to_edge_major
from .mesh import MeshBatch


class MeshUnpool(nn.Module):
    def __init__(self, unroll_target):
        super(MeshUnpool, self).__init__()
        self.unroll_target = unroll_target

    def __call__(self, features, meshes):
        return self.forward(features, meshes)

    def forward(self, features, meshes):
        """[B, C, E_lo] -> [B, C, unroll_target]; pops one pooling level off every mesh (mesh_unpool.py:30-41)."""
        return as_bce(self.forward_em(to_edge_major(features), MeshBatch.of(meshes, features.shape[2], features.device)))

    def forward_em(self, f_em, batch):
        ec_lo = batch.level().ec
        rec = batch.unroll()
        if f_em.shape[1] != rec.T:
            raise ValueError('MeshUnpool input has %d edge slots but the level being undone was pooled to %d'
                             % (f_em.shape[1], rec.T))
        if self.unroll_target < rec.E_in:
            raise ValueError('unroll_target %d is smaller than the restored level (%d)' % (self.unroll_target, rec.E_in))
        return UnpoolFn.apply(f_em, rec, batch.level().ec, ec_lo, self.unroll_target)
