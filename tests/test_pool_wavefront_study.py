"""Design study for a wavefront-parallel collapse ("next round"): how much of MeshPool's strictly ordered
edge-collapse sequence is really dependent?

The priorities never change after the sort (mesh_pool.py:184-192), so the candidate sequence of a level is static.  Here the
oracle's collapse loop runs with every array access of every pop recorded at row granularity (gemm/sides row, edges row,
vertex list, mask, v_mask, MeshUnion row); two pops are ordered only if one writes a row the other reads or writes.  The
length of the longest dependency chain against the number of pops bounds what any conflict-detecting parallel schedule can
gain, and a level-by-level schedule with P workers says what P collapse warps would get.  The global edge count (the loop
condition and the guards of __clean_side, mesh_pool.py:44,75-81) is left out of the dependencies: it only matters within
the last few collapses of a level, which a real implementation has to serialise.

The test also checks that the instrumented run is the oracle's run (same pops, outcomes, topology)."""
import numpy as np
import pytest
import torch

from meshcnn_b200 import synth
from oracle import meshcnn_oracle as O
from tests.helpers import prepared


class _TracedRun(O._PoolRun):
    def __init__(self, *a):
        super().__init__(*a)
        self.R, self.W = set(), set()

    def begin(self, e):
        self.R, self.W = {('mask', e)}, set()

    def has_boundaries(self, e):
        g = self.m.gemm_edges
        self.R.add(('g', e))
        for n in g[e]:
            if n != -1:
                self.R.add(('g', int(n)))
        return super().has_boundaries(e)

    def face_info(self, e, side):
        r = super().face_info(e, side)
        self.R.update({('g', e), ('g', r[0]), ('g', r[1])})
        return r

    def redirect(self, a, sa, b, sb):
        self.W.update({('g', int(a)), ('g', int(b))})
        super().redirect(a, sa, b, sb)

    def union(self, s, t):
        self.R.add(('grp', int(s)))
        self.W.add(('grp', int(t)))
        super().union(s, t)

    def remove_triplet(self, tri):
        common = set(self.m.edges[tri[0]].tolist())
        for k in tri:
            self.R.add(('ed', int(k)))
            self.W.add(('mask', int(k)))
            common &= set(self.m.edges[k].tolist())
        for v in common:
            self.W.add(('vm', int(v)))
        super().remove_triplet(tri)

    def one_ring_valid(self, e):
        m = self.m
        v0, v1 = int(m.edges[e, 0]), int(m.edges[e, 1])
        self.R.update({('ed', e), ('ve', v0), ('ve', v1)})
        for x in list(m.ve[v0]) + list(m.ve[v1]):
            self.R.add(('ed', int(x)))
        return super().one_ring_valid(e)

    def _remove_edge_trace(self, e):
        self.R.add(('ed', int(e)))
        for v in self.m.edges[e]:
            self.W.add(('ve', int(v)))

    def pool_side(self, e, side):
        ka, kb = int(self.m.gemm_edges[e, side]), int(self.m.gemm_edges[e, side + 1])
        self.W.add(('mask', kb))
        self._remove_edge_trace(kb)
        super().pool_side(e, side)

    def pool_edge(self, e):
        m = self.m
        v0, v1 = int(m.edges[e, 0]), int(m.edges[e, 1])
        out = super().pool_edge(e)
        if out == O.COLLAPSED:
            # merge_vertices (mesh.py:26-37) ran inside: its accesses, reconstructed (v0 / v1 were read before the call)
            self.R.add(('ed', e))
            self.W.update({('ve', v0), ('ve', v1), ('vm', v1), ('mask', e)})
            for r in np.nonzero((m.edges == v0).any(axis=1))[0]:       # rows that held v1 now hold v0 (a superset: all v0 rows)
                self.W.add(('ed', int(r)))
        return out


def _traced_pool(mesh, norms, target):
    """pool_mesh's loop (mesh_pool.py:41-56) with the traced run; returns the log and the per-pop access sets."""
    ec0 = mesh.edges_count
    order = O.collapse_order(norms, ec0)
    groups = O.SparseGroups(ec0)
    log = {'pops': [], 'outcomes': [], 'unions': [], 'triplets': []}
    run = _TracedRun(mesh, target, groups, log)
    acc, p = [], 0
    while mesh.edges_count > target:
        e = int(order[p]); p += 1
        run.begin(e)
        out = run.pool_edge(e) if run.mask[e] else O.SKIP_MASKED
        log['pops'].append(e)
        log['outcomes'].append(out)
        acc.append((out, frozenset(run.R), frozenset(run.W)))
    mesh.clean(run.mask, groups)
    return log, acc


def _levels(acc):
    """Earliest level of every pop under read/write dependencies (row granularity)."""
    lw, lr, lev = {}, {}, []
    for out, R, W in acc:
        l = 0
        for k in R | W:
            l = max(l, lw.get(k, -1) + 1)
        for k in W:
            l = max(l, lr.get(k, -1) + 1)
        lev.append(l)
        for k in R:
            lr[k] = max(lr.get(k, -1), l)
        for k in W:
            lw[k] = l
    return np.array(lev)


def _study(seeds, chain):
    rows = []
    for s in seeds:
        vs, faces = synth.make('ico', s, nu=5)
        data = prepared(vs, faces)
        ref, tr = O.OracleMesh(data), O.OracleMesh(data)
        rs = np.random.RandomState(100 + s)
        for T in chain:
            ec = ref.edges_count
            norms = rs.rand(ec).astype(np.float32)
            f = torch.zeros(1, ec)
            _, lg_ref = O.pool_mesh(ref, f, T, norms=norms)
            lg, acc = _traced_pool(tr, norms, T)
            assert lg['pops'] == lg_ref['pops'] and lg['outcomes'] == lg_ref['outcomes']
            np.testing.assert_array_equal(ref.gemm_edges, tr.gemm_edges)
            np.testing.assert_array_equal(ref.edges, tr.edges)
            assert ref.ve == tr.ve
            lev = _levels(acc)
            work = np.array([0.0 if o == O.SKIP_MASKED else 1.0 for o, _, _ in acc])      # skipped pops are free
            n = int(work.sum())
            depth = int(lev.max()) + 1
            sched = {P: int(sum(-(-int(work[lev == l].sum()) // P) for l in range(depth))) for P in (2, 4, 8)}
            rows.append((s, ec, T, len(acc), n, depth, sched))
    return rows


def test_collapse_sequence_has_exploitable_parallelism():
    rows = _study(seeds=range(6), chain=[600, 450, 300, 210])
    print()
    for T in (600, 450, 300, 210):
        sel = [r for r in rows if r[2] == T]
        n = np.mean([r[4] for r in sel]); d = np.mean([r[5] for r in sel])
        sp = {P: np.mean([r[4] / max(1, r[6][P]) for r in sel]) for P in (2, 4, 8)}
        print('  -> %3d edges: %5.1f examined candidates per mesh, longest dependency chain %5.1f (%.1fx), '
              'level schedule with 2 / 4 / 8 workers: %.2fx / %.2fx / %.2fx' % (T, n, d, n / d, sp[2], sp[4], sp[8]))
    for r in rows:
        assert r[5] <= r[4] or r[4] == 0
    assert np.mean([r[4] / r[5] for r in rows]) > 1.5            # the strict order is far from a single chain
