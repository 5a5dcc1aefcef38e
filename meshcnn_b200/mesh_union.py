"""Drop-in for models/layers/mesh_union.py (host side).

Inside MeshPool the union bookkeeping runs in the collapse kernel (the pour log -> CSR + CSC + occurrences); this class is
the host object with the reference's surface -- ``union / remove_group / get_group / get_occurrences / get_groups /
rebuild_features(_average) / prepare_groups`` and the ``groups`` attribute -- for code that drives a MeshUnion itself
(``Mesh.history_data['collapses']``, mesh.py:161-162; a host-side pool loop; inspection).  Same results as the reference's
dense n x n fp32 matrix, kept as sparse rows {member: multiplicity} (> 99.8 % zeros) until someone reads or
assigns ``groups``, at which point it becomes the dense tensor for good (in-place edits by the caller must stick)."""
import numpy as np
import torch


class MeshUnion:
    def __init__(self, n, device=torch.device('cpu')):
        self.__size = n
        self.device = device
        self.rebuild_features = self.rebuild_features_average
        self._rows = [{i: 1} for i in range(n)]                  # identity (mesh_union.py:9), sparse
        self._dense = None

    # -- the `groups` attribute --------------------------------------------------------------------------------
    @property
    def groups(self):
        if self._dense is None:
            g = torch.zeros(len(self._rows), self.__size)
            for i, r in enumerate(self._rows):
                if r:
                    g[i, list(r.keys())] = torch.tensor(list(r.values()), dtype=torch.float32)
            self._dense, self._rows = g.to(self.device), None
        return self._dense

    @groups.setter
    def groups(self, value):
        self._dense, self._rows = value, None

    @property
    def rows(self):
        """Sparse rows (None once the dense view has been taken)."""
        return self._rows

    # -- reference API -------------------------------------------------------------------------------------------
    def union(self, source, target):                              # mesh_union.py:11-12
        if self._rows is None:
            self._dense[target, :] += self._dense[source, :]
            return
        rt = self._rows[target]
        for m, c in list(self._rows[source].items()):
            rt[m] = rt.get(m, 0) + c

    def remove_group(self, index):                                # mesh_union.py:14-15
        return

    def get_group(self, edge_key):                                # mesh_union.py:17-18
        return self.groups[edge_key, :]

    def get_occurrences(self):                                    # mesh_union.py:20-21
        if self._rows is None:
            return torch.sum(self._dense, 0)
        occ = np.zeros(self.__size, dtype=np.float32)
        for r in self._rows:
            for m, c in r.items():
                occ[m] += c
        return torch.from_numpy(occ).to(self.device)

    def get_groups(self, tensor_mask):                            # mesh_union.py:23-25 (clamps ALL rows, then selects)
        self.groups = torch.clamp(self.groups, 0, 1)
        return self.groups[tensor_mask, :]

    def prepare_groups(self, features, mask):                     # mesh_union.py:38-44
        """groups <- clamp(groups[mask], 0, 1)^T zero-padded to features.shape[1] rows: [E, n_surviving]."""
        keep = np.nonzero(np.asarray(mask, dtype=bool))[0]
        E = features.shape[1]
        if self._rows is not None:
            g = torch.zeros(max(E, self.__size), len(keep))
            for j, i in enumerate(keep):
                r = self._rows[int(i)]
                if r:
                    g[list(r.keys()), j] = 1.0
            self.groups = g.to(self.device)
            return
        g = torch.clamp(self._dense[torch.from_numpy(keep)], 0, 1).transpose(1, 0)
        pad = E - g.shape[0]
        if pad > 0:
            g = torch.cat([g, torch.zeros(pad, g.shape[1], dtype=g.dtype, device=g.device)], 0)
        self.groups = g

    def rebuild_features_average(self, features, mask, target_edges):   # mesh_union.py:27-36
        """features [C, E(, 1)] -> [C, target_edges]: mean over each surviving edge's (clamped) group, zero-padded."""
        self.prepare_groups(features, mask)
        g = self.groups.to(features.device)
        f = features.squeeze(-1)
        fe = torch.matmul(f, g)
        fe = fe / torch.sum(g, 0).expand(fe.shape)
        pad = target_edges - fe.shape[1]
        if pad > 0:
            fe = torch.cat([fe, torch.zeros(fe.shape[0], pad, dtype=fe.dtype, device=fe.device)], 1)
        return fe
