"""CPU oracle for the MeshCNN per-edge hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file.  The product package (meshcnn_b200/) never does; it fails loudly when its CUDA library is missing.

What this is: an independent restatement, in NumPy / PyTorch-CPU, of the algorithm the reference
(ranahanocka/MeshCNN, pure Python) runs on the path MeshConv -> MeshPool (+MeshUnion) -> MeshUnpool.
Every function cites the reference file:line it follows.  It is deliberately written with different data
structures from the CUDA kernels (Python lists / dicts here; union-find, linked lists and CSR there), so
that agreement between the two is evidence and not a tautology.

Parity pinning: the reference's own tests hold no layer-level golden vectors, so the pins are
outputs of the reference itself, generated in the build container by tools/make_golden.py (which imports
/root/reference unmodified) and committed under tests/golden/.  tests/test_oracle_golden.py checks this
oracle against every one of them: collapse order, per-pop outcome, union log, mask, post-clean
gemm_edges / sides / edges / ve / v_mask / edges_count bit-exactly; groups and occurrences exactly; features
and gradients to 1e-5.

Group bookkeeping comes in two flavours with identical results:
  * ``SparseGroups``  row -> {member: multiplicity} dicts  (fast; usable at 170k edges)
  * ``DenseGroups``   the reference's n x n fp32 matrix with row adds (mesh_union.py:9-12); this is what
                      the CPU baseline times, because it is what the reference executes.
"""
import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------
# mesh state
# ----------------------------------------------------------------------------------------------------------
class OracleMesh:
    """Host topology state of one mesh (mesh.py:10-21) built from prepared arrays."""

    def __init__(self, data, hold_history=False):
        self.vs = np.array(data.vs, dtype=np.float64)
        self.v_mask = np.array(data.v_mask, dtype=bool)
        self.edges = np.array(data.edges, dtype=np.int32)
        self.gemm_edges = np.array(data.gemm_edges, dtype=np.int64)
        self.sides = np.array(data.sides, dtype=np.int64)
        self.ve = [list(l) for l in data.ve]
        self.edges_count = int(data.edges_count)
        self.features = data.features
        self.edge_areas = data.edge_areas
        self.filename = getattr(data, 'filename', 'unknown')
        self.pool_count = 0
        self.history = None
        if hold_history:                                      # mesh.py:151-162 (only the unpool-relevant part)
            self.history = {'groups': [], 'occurrences': [], 'gemm_edges': [self.gemm_edges.copy()],
                            'edges_count': [self.edges_count]}

    # mesh.py:42-48 -- first-occurrence removal from both endpoint lists; raises ValueError when absent
    def remove_edge(self, e):
        for v in self.edges[e]:
            self.ve[v].remove(e)

    # mesh.py:26-37
    def merge_vertices(self, e):
        self.remove_edge(e)
        v0, v1 = int(self.edges[e, 0]), int(self.edges[e, 1])
        self.vs[v0] = (self.vs[v0] + self.vs[v1]) / 2
        self.v_mask[v1] = False
        hit = self.edges == v1                     # every row: live, dead, stale-reachable
        self.ve[v0].extend(self.ve[v1])
        self.edges[hit] = v0

    # mesh.py:50-71
    def clean(self, mask, groups):
        mask = mask.astype(bool)
        self.gemm_edges = self.gemm_edges[mask]
        self.edges = self.edges[mask]
        self.sides = self.sides[mask]
        remap = np.zeros(len(mask) + 1, dtype=np.int64)        # dead -> 0; slot [-1] keeps -1
        remap[-1] = -1
        remap[:-1][mask] = np.arange(int(mask.sum()))
        self.gemm_edges = remap[self.gemm_edges]
        self.ve = [[int(remap[e]) for e in l] for l in self.ve]
        if self.history is not None:                          # mesh.py:182-192: occurrences BEFORE clamping (F7)
            self.history['occurrences'].append(groups.occurrences())
            self.history['groups'].append(groups.clamped_rows(mask))
            self.history['gemm_edges'].append(self.gemm_edges.copy())
            self.history['edges_count'].append(self.edges_count)
        self.pool_count += 1

    # mesh.py:194-198 -- only gemm_edges and edges_count are restored
    def unroll_gemm(self):
        self.history['gemm_edges'].pop()
        self.gemm_edges = self.history['gemm_edges'][-1]
        self.history['edges_count'].pop()
        self.edges_count = self.history['edges_count'][-1]


# ----------------------------------------------------------------------------------------------------------
# MeshUnion restatements (mesh_union.py)
# ----------------------------------------------------------------------------------------------------------
class SparseGroups:
    def __init__(self, n):
        self.n = n
        self.rows = [{i: 1} for i in range(n)]                 # identity, mesh_union.py:9

    def union(self, s, t):                                    # mesh_union.py:11-12 (row add keeps multiplicity)
        rt = self.rows[t]
        for m, c in list(self.rows[s].items()):
            rt[m] = rt.get(m, 0) + c

    def occurrences(self):                                    # mesh_union.py:20-21, all rows, un-clamped
        occ = np.zeros(self.n, dtype=np.float32)
        for r in self.rows:
            for m, c in r.items():
                occ[m] += c
        return occ

    def clamped_rows(self, mask):                             # mesh_union.py:23-25 -> list of sorted member arrays
        return [np.array(sorted(self.rows[i].keys()), dtype=np.int64) for i in np.nonzero(mask)[0]]


class DenseGroups:
    """The reference's own representation: n x n fp32, dense row adds."""

    def __init__(self, n):
        self.n = n
        self.g = torch.eye(n)

    def union(self, s, t):
        self.g[t, :] += self.g[s, :]

    def occurrences(self):
        return torch.sum(self.g, 0).numpy().copy()

    def clamped_rows(self, mask):
        g = torch.clamp(self.g, 0, 1)[torch.from_numpy(np.asarray(mask, dtype=bool))]
        return [np.nonzero(r)[0].astype(np.int64) for r in g.numpy()]


def segment_mean(features, rows, target):
    """rebuild_features_average (mesh_union.py:27-36): features [C, E] torch -> [C, target]; differentiable."""
    C, E = features.shape
    n = len(rows)
    G = torch.zeros(E, n, dtype=features.dtype)
    for r, members in enumerate(rows):
        G[members, r] = 1
    out = (features @ G) / G.sum(0)
    if target > n:
        out = torch.nn.functional.pad(out, (0, target - n))
    return out


# ----------------------------------------------------------------------------------------------------------
# MeshPool (mesh_pool.py)
# ----------------------------------------------------------------------------------------------------------
# per-pop outcome codes written to the log
SKIP_MASKED, REJ_BOUNDARY, REJ_SIDE0, REJ_SIDE2, REJ_ONE_RING, COLLAPSED = 0, 1, 2, 3, 4, 5


class PoolExhausted(IndexError):
    """heappop from an empty queue (mesh_pool.py:49-50)."""


def _other_side(s):                                            # mesh_pool.py:156-158
    return s + 1 - 2 * (s % 2)


class _PoolRun:
    def __init__(self, mesh, target, groups, log):
        self.m, self.T, self.G, self.log = mesh, target, groups, log
        self.mask = np.ones(mesh.edges_count, dtype=bool)

    def union(self, s, t):
        self.G.union(int(s), int(t))
        self.log['unions'].append((int(s), int(t)))

    def has_boundaries(self, e):                               # mesh_pool.py:87-92
        g = self.m.gemm_edges
        for n in g[e]:
            if n == -1 or -1 in g[n]:
                return True
        return False

    def face_info(self, e, side):                              # mesh_pool.py:160-170
        g, s = self.m.gemm_edges, self.m.sides
        ka, kb = int(g[e, side]), int(g[e, side + 1])
        sa, sb = int(s[e, side]), int(s[e, side + 1])
        osa = (sa - sa % 2 + 2) % 4
        osb = (sb - sb % 2 + 2) % 4
        oka = [int(g[ka, osa]), int(g[ka, osa + 1])]
        okb = [int(g[kb, osb]), int(g[kb, osb + 1])]
        return ka, kb, sa, sb, osa, osb, oka, okb

    def redirect(self, a, sa, b, sb):                          # mesh_pool.py:140-145
        g, s = self.m.gemm_edges, self.m.sides
        g[a, sa] = b
        g[b, sb] = a
        s[a, sa] = sb
        s[b, sb] = sa

    def get_invalids(self, e, side):                           # mesh_pool.py:115-138 (mutates)
        ka, kb, sa, sb, osa, osb, oka, okb = self.face_info(e, side)
        sh = []
        for i in range(2):                                     # mesh_pool.py:147-154
            for j in range(2):
                if oka[i] == okb[j]:
                    sh.extend([i, j])
        if not sh:
            return []
        assert len(sh) == 2
        s = self.m.sides
        mid = oka[sh[0]]
        ua, ub = oka[1 - sh[0]], okb[1 - sh[1]]
        usa, usb = int(s[ka, osa + 1 - sh[0]]), int(s[kb, osb + 1 - sh[1]])
        self.redirect(e, side, ua, usa)
        self.redirect(e, side + 1, ub, usb)
        self.redirect(ua, _other_side(usa), ub, _other_side(usb))
        self.union(ka, e); self.union(kb, e)
        self.union(ka, ua); self.union(mid, ua)
        self.union(kb, ub); self.union(mid, ub)
        return [ka, kb, mid]

    def remove_triplet(self, tri):                             # mesh_pool.py:172-182 (no remove_edge: F9)
        common = set(self.m.edges[tri[0]].tolist())
        for k in tri:
            common &= set(self.m.edges[k].tolist())
            self.mask[k] = False
        self.m.edges_count -= 3
        assert len(common) == 1
        self.m.v_mask[common.pop()] = False
        self.log['triplets'].append(tuple(int(k) for k in tri))

    def clean_side(self, e, side):                             # mesh_pool.py:74-85
        m = self.m
        if m.edges_count <= self.T:
            return False
        inv = self.get_invalids(e, side)
        while inv and m.edges_count > self.T:
            self.remove_triplet(inv)
            if m.edges_count <= self.T:
                return False
            if self.has_boundaries(e):
                return False
            inv = self.get_invalids(e, side)
        return True

    def one_ring_valid(self, e):                               # mesh_pool.py:95-100 (reads stale ve entries: F9)
        m = self.m
        v0, v1 = int(m.edges[e, 0]), int(m.edges[e, 1])
        a = set(m.edges[m.ve[v0]].reshape(-1).tolist())
        b = set(m.edges[m.ve[v1]].reshape(-1).tolist())
        return len(a & (b - {v0, v1})) == 2

    def pool_side(self, e, side):                              # mesh_pool.py:102-113
        ka, kb, sa, sb, _, osb, _, okb = self.face_info(e, side)
        s = self.m.sides
        self.redirect(ka, sa - sa % 2, okb[0], int(s[kb, osb]))
        self.redirect(ka, sa - sa % 2 + 1, okb[1], int(s[kb, osb + 1]))
        self.union(kb, ka)
        self.union(e, ka)
        self.mask[kb] = False
        self.m.remove_edge(kb)
        self.m.edges_count -= 1

    def pool_edge(self, e):                                    # mesh_pool.py:58-72
        if self.has_boundaries(e):
            return REJ_BOUNDARY
        if not self.clean_side(e, 0):
            return REJ_SIDE0
        if not self.clean_side(e, 2):
            return REJ_SIDE2
        if not self.one_ring_valid(e):
            return REJ_ONE_RING
        self.pool_side(e, 0)
        self.pool_side(e, 2)
        self.m.merge_vertices(e)
        self.mask[e] = False
        self.m.edges_count -= 1
        return COLLAPSED


def collapse_order(norms_fp32, ec):
    """heapify + heappop over a fixed set == ascending sort by (norm as double, edge id) (mesh_pool.py:184-192, F2)."""
    n = np.asarray(norms_fp32[:ec], dtype=np.float32).astype(np.float64)
    return np.lexsort((np.arange(ec), n))


def squared_norms(features):
    """torch.sum(f * f, 0) in fp32 (mesh_pool.py:186); features [C, ec] torch tensor."""
    return torch.sum(features * features, 0)


def pool_mesh(mesh, features, target, norms=None, dense=False):
    """__pool_main for one mesh (mesh_pool.py:41-56).  features: [C, E] torch (E >= edges_count).
    Returns (out [C, target], log).  ``norms`` (fp32, length >= ec) may be supplied to teacher-force the order."""
    ec0 = mesh.edges_count
    if norms is None:
        norms = squared_norms(features[:, :ec0].detach().float()).numpy()
    order = collapse_order(norms, ec0)
    groups = DenseGroups(ec0) if dense else SparseGroups(ec0)
    log = {'pops': [], 'outcomes': [], 'unions': [], 'triplets': [], 'norms': np.asarray(norms[:ec0], dtype=np.float32)}
    run = _PoolRun(mesh, target, groups, log)
    p = 0
    while mesh.edges_count > target:
        if p == ec0:
            raise PoolExhausted('pop from an empty queue')
        e = int(order[p]); p += 1
        log['pops'].append(e)
        log['outcomes'].append(run.pool_edge(e) if run.mask[e] else SKIP_MASKED)
    mask = run.mask
    mesh.clean(mask, groups)                                   # BEFORE the rebuild (F7)
    rows = groups.clamped_rows(mask)
    out = segment_mean(features.reshape(features.shape[0], -1), rows, target)
    log['mask'] = mask.copy()
    log['rows'] = rows
    log['occurrences'] = groups.occurrences()
    return out, log


def pool_forward(x, meshes, target, norms=None, dense=False):
    """MeshPool.forward (mesh_pool.py:23-39): x [B, C, E] or [B, C, E, 1] -> [B, C, target] (+ per-mesh logs)."""
    if x.dim() == 4:
        x = x.squeeze(-1)
    outs, logs = [], []
    for b, m in enumerate(meshes):
        o, lg = pool_mesh(m, x[b], target, None if norms is None else norms[b], dense)
        outs.append(o)
        logs.append(lg)
    return torch.stack(outs, 0), logs


# ----------------------------------------------------------------------------------------------------------
# MeshUnpool (mesh_unpool.py:30-41)
# ----------------------------------------------------------------------------------------------------------
def unpool_forward(x, meshes, unroll_target):
    """x [B, C, E_lo] -> [B, C, unroll_target]; pops one history level off every mesh."""
    B, C, E_lo = x.shape
    mats = []
    for m in meshes:
        rows = m.history['groups'].pop()
        occ = m.history['occurrences'].pop()
        U = torch.zeros(E_lo, unroll_target, dtype=x.dtype)
        for r, members in enumerate(rows):
            U[r, members] = 1
        o = torch.ones(unroll_target, dtype=x.dtype)          # occurrences padded with 1 (mesh_unpool.py:23-28)
        o[:len(occ)] = torch.from_numpy(np.asarray(occ)).to(x.dtype)
        mats.append(U / o)
        m.unroll_gemm()
    return torch.matmul(x, torch.stack(mats, 0))


# ----------------------------------------------------------------------------------------------------------
# MeshConv (mesh_conv.py)
# ----------------------------------------------------------------------------------------------------------
def neighbour_table(meshes, E):
    """[B, E, 5] int64: self id + 4 one-ring neighbours; -1 = zero feature; slots >= edges_count all point at
    edge 0 (pad value 0 is applied before the +1 shift: mesh_conv.py:82 then :50)."""
    B = len(meshes)
    t = np.zeros((B, E, 5), dtype=np.int64)
    for b, m in enumerate(meshes):
        ec = m.edges_count
        t[b, :ec, 0] = np.arange(ec)
        t[b, :ec, 1:] = m.gemm_edges[:ec]
    return t


def conv_forward(x, meshes, weight, bias=None):
    """MeshConv.forward (mesh_conv.py:20-26, :39-70).  x [B, Cin, E] (or 4-D), weight [Cout, Cin, 1, 5].
    Returns [B, Cout, E, 1].  Differentiable (autograd supplies the reference's backward)."""
    if x.dim() == 4:
        x = x.squeeze(-1)
    B, C, E = x.shape
    nb = torch.from_numpy(neighbour_table(meshes, E)) + 1                  # 0 -> the zero pad column
    xp = torch.cat([x.new_zeros(B, C, 1), x], 2)                           # mesh_conv.py:47-49
    idx = nb.reshape(B, 1, E * 5).expand(B, C, E * 5)
    f = torch.gather(xp, 2, idx).reshape(B, C, E, 5)
    g = torch.stack([f[..., 0], f[..., 1] + f[..., 3], f[..., 2] + f[..., 4],
                     (f[..., 1] - f[..., 3]).abs(), (f[..., 2] - f[..., 4]).abs()], 3)   # mesh_conv.py:65-69
    return torch.nn.functional.conv2d(g, weight, bias)                     # mesh_conv.py:25


def conv_forward_reference_cost(x, meshes, weight, bias=None):
    """Same result as conv_forward, but following the reference's op sequence (per-mesh pad_gemm, float index
    arithmetic, permute+contiguous, index_select; mesh_conv.py:28-84) so that timing it is a fair CPU baseline."""
    if x.dim() == 4:
        x = x.squeeze(-1)
    B, C, E = x.shape
    tabs = []
    for m in meshes:
        t = torch.tensor(m.gemm_edges[:m.edges_count]).float()
        t = torch.cat((torch.arange(m.edges_count).float().unsqueeze(1), t), 1)
        tabs.append(torch.nn.functional.pad(t, (0, 0, 0, E - m.edges_count)).unsqueeze(0))
    gi = torch.cat(tabs, 0) + 1
    xp = torch.cat((x.new_zeros(B, C, 1), x), 2)
    offs = (torch.arange(B).float() * (E + 1)).view(B, 1, 1)
    flat = (gi + offs).view(-1).long()
    rows = xp.permute(0, 2, 1).contiguous().view(B * (E + 1), C)
    f = torch.index_select(rows, 0, flat).view(B, E, 5, C).permute(0, 3, 1, 2)
    g = torch.stack([f[..., 0], f[..., 1] + f[..., 3], f[..., 2] + f[..., 4],
                     torch.abs(f[..., 1] - f[..., 3]), torch.abs(f[..., 2] - f[..., 4])], 3)
    return torch.nn.functional.conv2d(g, weight, bias)
