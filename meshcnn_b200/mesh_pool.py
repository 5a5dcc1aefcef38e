"""Drop-in for models/layers/mesh_pool.py.  The heapq-driven Python collapse loop becomes: one norm kernel, one
collapse kernel (one CTA per mesh: sort + sequential collapse + Mesh.clean + MeshUnion export) and one
segment-mean kernel.  Collapse order, failed-attempt side effects and the rebuilt topology are bit-exact with the
reference for the same fp32 norms (tests/test_pool_gpu.py)."""
import torch
import torch.nn as nn

from .functional import SegMeanFn, pool_norms
from .layout import as_bce, to_edge_major
from .mesh import MeshBatch


class MeshPool(nn.Module):

    def __init__(self, target, multi_thread=False):
        super(MeshPool, self).__init__()
        self.__out_target = target
        self.__multi_thread = multi_thread      # accepted for API parity; every mesh already has its own CTA
        self.norm_override = None               # tests: feed the oracle's fp32 norms ([B, E]) to pin the order
        self.last_record = None

    def __call__(self, fe, meshes):
        return self.forward(fe, meshes)

    @property
    def target(self):
        return self.__out_target

    def forward(self, fe, meshes):
        """fe [B, C, E] or [B, C, E, 1] -> [B, C, target]; every Mesh is pooled in place (device state)."""
        return as_bce(self.forward_em(to_edge_major(fe), MeshBatch.of(meshes, fe.shape[2], fe.device)))

    def forward_em(self, x_em, batch):
        lv = batch.level(x_em.shape[1])
        if lv is not batch.level():
            raise ValueError('MeshPool input has %d edge slots, the mesh level has %d' % (x_em.shape[1], batch.level().E))
        with torch.no_grad():
            norms = pool_norms(x_em.detach().contiguous()) if self.norm_override is None else \
                self.norm_override.to(device=x_em.device, dtype=torch.float32).contiguous()
        ec_in = lv.ec
        rec = batch.pool(norms, self.__out_target)
        self.last_record = rec
        return SegMeanFn.apply(x_em, rec, ec_in, batch.level().ec)

    @staticmethod
    def has_boundaries(mesh, edge_id):
        """Public in the reference (mesh_pool.py:87-92); host-side helper on the mirrored arrays."""
        g = mesh.gemm_edges
        for edge in g[edge_id]:
            if edge == -1 or -1 in g[edge]:
                return True
        return False
