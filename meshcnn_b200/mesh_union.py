"""Drop-in for models/layers/mesh_union.py.

Inside MeshPool the union bookkeeping runs in the collapse kernel (sparse rows with multiplicities, emitted as
CSR + CSC + occurrences); this class is the host-visible view of one mesh's result with the reference's method
names, for code that inspects groups (e.g. ``Mesh.history_data['collapses']``, mesh.py:161-162)."""
import numpy as np
import torch


class MeshUnion:
    def __init__(self, n, device=torch.device('cpu')):
        self.__size = n
        self.device = device
        self.rebuild_features = self.rebuild_features_average
        self.rows = [{i: 1} for i in range(n)]                  # identity (mesh_union.py:9), sparse

    @property
    def groups(self):
        """Dense view, materialised on demand only (n x n fp32 like the reference attribute)."""
        g = torch.zeros(self.__size, self.__size)
        for i, r in enumerate(self.rows):
            for m, c in r.items():
                g[i, m] = c
        return g.to(self.device)

    def union(self, source, target):                              # mesh_union.py:11-12
        rt = self.rows[target]
        for m, c in list(self.rows[source].items()):
            rt[m] = rt.get(m, 0) + c

    def remove_group(self, index):                                # mesh_union.py:14-15
        return

    def get_group(self, edge_key):                                # mesh_union.py:17-18
        return self.groups[edge_key, :]

    def get_occurrences(self):                                    # mesh_union.py:20-21
        occ = np.zeros(self.__size, dtype=np.float32)
        for r in self.rows:
            for m, c in r.items():
                occ[m] += c
        return torch.from_numpy(occ).to(self.device)

    def get_groups(self, tensor_mask):                            # mesh_union.py:23-25
        g = torch.clamp(self.groups, 0, 1)
        return g[tensor_mask, :]

    def rebuild_features_average(self, features, mask, target_edges):
        raise RuntimeError('feature rebuilding runs on the device inside meshcnn_b200.MeshPool (mcnn_segmean_fwd); '
                           'the host MeshUnion is an inspection view only')

    def prepare_groups(self, features, mask):
        raise RuntimeError('see rebuild_features_average')
