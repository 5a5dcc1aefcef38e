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
        self.rewired = False                                     # __get_invalids found a valence-3 vertex and rewired (mesh_pool.py:115-138)

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

    def get_invalids(self, e, side):
        inv = super().get_invalids(e, side)
        if inv:
            self.rewired = True
        return inv

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


def _ball(edges, mask, seeds, radius):
    """Vertices within `radius` edge hops of `seeds` over the live edges; returns {vertex: distance}."""
    adj = {}
    for a, b in edges[mask]:
        adj.setdefault(int(a), []).append(int(b))
        adj.setdefault(int(b), []).append(int(a))
    dist = {v: 0 for v in seeds}
    frontier = list(seeds)
    for d in range(1, radius + 1):
        nxt = []
        for v in frontier:
            for w in adj.get(v, ()):
                if w not in dist:
                    dist[w] = d
                    nxt.append(w)
        frontier = nxt
    return dist


def _traced_pool(mesh, norms, target, radius=4):
    """pool_mesh's loop (mesh_pool.py:41-56) with the traced run; returns the log and, per pop, (outcome, read rows, written
    rows, ball of `radius` around the candidate's endpoints at pop start, touched vertices with their distance)."""
    ec0 = mesh.edges_count
    order = O.collapse_order(norms, ec0)
    groups = O.SparseGroups(ec0)
    log = {'pops': [], 'outcomes': [], 'unions': [], 'triplets': []}
    run = _TracedRun(mesh, target, groups, log)
    acc, p = [], 0
    while mesh.edges_count > target:
        e = int(order[p]); p += 1
        run.begin(e)
        if not run.mask[e]:
            out, ball, far = O.SKIP_MASKED, {}, 0
        else:
            ed0 = mesh.edges.copy()
            mesh_vmask0 = mesh.v_mask.copy()
            ball = _ball(ed0, run.mask, [int(ed0[e, 0]), int(ed0[e, 1])], radius)
            out = run.pool_edge(e)
            touched = set()
            for kind, i in run.R | run.W:
                touched.update([int(i)] if kind in ('ve', 'vm') else [int(ed0[i, 0]), int(ed0[i, 1])])
            live = [v for v in touched if mesh_vmask0[v]]             # rows of dead edges may still name dead vertices (F9)
            far = max([ball.get(v, radius + 1) for v in live] + [0])  # radius + 1: further away than the ball
        log['pops'].append(e)
        log['outcomes'].append(out)
        acc.append((out, frozenset(run.R), frozenset(run.W), ball, far, run.rewired))
    mesh.clean(run.mask, groups)
    return log, acc


def _levels(acc):
    """Earliest level of every pop under read/write dependencies (row granularity)."""
    lw, lr, lev = {}, {}, []
    for out, R, W, _, _, _ in acc:
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


def _levels_by_ball(acc, r):
    """Earliest level of every pop when two pops are ordered whenever their radius-r vertex balls intersect -- a rule a kernel
    can evaluate BEFORE running a candidate (stamp the ball, look for foreign stamps)."""
    last, lev = {}, []
    for out, _, _, ball, _, _ in acc:
        if out == O.SKIP_MASKED:
            lev.append(0)
            continue
        vs_ = [v for v, d in ball.items() if d <= r]
        l = max([last.get(v, -1) + 1 for v in vs_] + [0])
        lev.append(l)
        for v in vs_:
            last[v] = l
    return np.array(lev)


def _rule_misses(acc, r):
    """Pairs of pops with a true row conflict whose radius-r balls do NOT intersect (the rule would run them concurrently)."""
    touch = {}
    for j, (out, R, W, ball, _, _) in enumerate(acc):
        if out == O.SKIP_MASKED:
            continue
        for k in R | W:
            touch.setdefault(k, []).append((j, k in W))
    balls = [set(v for v, d in a[3].items() if d <= r) for a in acc]
    miss = set()
    for k, lst in touch.items():
        for x in range(len(lst)):
            for y in range(x + 1, len(lst)):
                (i, wi), (j, wj) = lst[x], lst[y]
                if (wi or wj) and not (balls[i] & balls[j]):
                    miss.add((i, j))
    return len(miss)


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
            work = np.array([0.0 if a[0] == O.SKIP_MASKED else 1.0 for a in acc])      # skipped pops are free
            n = int(work.sum())
            depth = int(lev.max()) + 1
            sched = {P: int(sum(-(-int(work[lev == l].sum()) // P) for l in range(depth))) for P in (2, 4, 8)}
            far = max(a[4] for a in acc)
            ball = {}
            for r in (2, 3):
                lb = _levels_by_ball(acc, r)
                ball[r] = (int(lb.max()) + 1, int(sum(-(-int(work[lb == l].sum()) // 4) for l in range(int(lb.max()) + 1))), _rule_misses(acc, r))
            rows.append((s, ec, T, len(acc), n, depth, sched, far, ball, sum(1 for a in acc if a[5])))
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
    print('  candidates whose __get_invalids rewires (they cannot be inspected read-only and would run alone): %d of %d'
          % (sum(r[9] for r in rows), sum(r[4] for r in rows)))
    far = max(r[7] for r in rows)
    print('  every LIVE vertex whose rows a pop touches lies within %d hops of the candidate edge' % far)
    for rr in (2, 3):
        d = np.mean([r[8][rr][0] for r in rows]); n = np.mean([r[4] for r in rows])
        sp4 = np.mean([r[4] / max(1, r[8][rr][1]) for r in rows])
        miss = sum(r[8][rr][2] for r in rows)
        print('  a-priori rule "radius-%d vertex balls must be disjoint": chain %.1f of %.1f pops, 4 workers %.2fx, '
              'true conflicts the rule misses: %d' % (rr, d, n, sp4, miss))
    for r in rows:
        assert r[5] <= r[4] or r[4] == 0
    assert np.mean([r[4] / r[5] for r in rows]) > 1.5            # the strict order is far from a single chain


# ---------------------------------------------------------------------------------------------------------------------
# Deterministic reservations (inspect -> reserve -> commit rounds), simulated on the oracle
# ---------------------------------------------------------------------------------------------------------------------
import copy


def _reservation_pool(mesh, norms, target, window, rmax=15):
    """The collapse loop as rounds over the first `window` pending candidates: (1) inspect -- a dry run of each candidate on a
    copy of the current state gives the rows it would touch; (2) reserve -- every row goes to the earliest candidate that
    wants it; (3) commit -- in candidate order, a candidate runs for real iff it holds all its rows.  The edge count is handled
    conservatively: a candidate may run ahead of k earlier, still pending ones only while count - rmax * (k + 1) > target
    (each collapse removes 3 edges plus 3 per removed triplet), so the loop condition and the guards of __clean_side see
    what they would see in the strict order.  Returns (log of the committed order, rounds, examined candidates)."""
    ec0 = mesh.edges_count
    order = [int(e) for e in O.collapse_order(norms, ec0)]
    groups = O.SparseGroups(ec0)
    log = {'pops': [], 'outcomes': [], 'unions': [], 'triplets': []}
    run = _TracedRun(mesh, target, groups, log)
    pending, rounds, examined = list(order), 0, 0
    while mesh.edges_count > target:
        pending = [e for e in pending if run.mask[e]]            # masked candidates are skipped when popped (mesh_pool.py:51)
        assert pending, 'queue exhausted'
        win = pending[:window]
        foot = []
        for e in win:                                            # inspect: dry run on a copy of the committed state
            trial = copy.deepcopy(run)
            trial.begin(e)
            trial.pool_edge(e)
            foot.append(trial.R | trial.W)
        owner = {}
        for pos, fp in enumerate(foot):                          # reserve: earliest candidate wins every row
            for k in fp:
                owner.setdefault(k, pos)
        done = []
        for pos, e in enumerate(win):                            # commit in candidate order
            held = all(owner[k] == pos for k in foot[pos])
            ahead = pos - len(done)                              # earlier candidates of the window that stay pending
            if not held or mesh.edges_count - rmax * (ahead + 1) <= target:
                if pos == 0:                                     # the head of the queue always runs: strict order from here
                    held = True
                else:
                    continue
            if mesh.edges_count <= target:
                break
            run.begin(e)
            out = run.pool_edge(e)
            assert (run.R | run.W) == foot[pos] or pos == 0, 'footprint changed between inspection and commit'
            log['pops'].append(e)
            log['outcomes'].append(out)
            done.append(e)
            examined += 1
        assert done
        pending = [e for e in pending if e not in set(done)]
        rounds += 1
    mesh.clean(run.mask, groups)
    return log, rounds, examined, run.mask, groups


@pytest.mark.parametrize('window', [4, 8])
def test_reservation_rounds_reproduce_the_strict_order(window):
    """The wavefront schedule must leave exactly the topology, mask and groups of the strict one-at-a-time order."""
    tot_rounds = tot_exam = 0
    for s in range(4):
        vs, faces = synth.make('ico', s, nu=5)
        data = prepared(vs, faces)
        ref, par = O.OracleMesh(data), O.OracleMesh(data)
        rs = np.random.RandomState(200 + s)
        for T in [600, 450, 300, 210]:
            ec = ref.edges_count
            norms = rs.rand(ec).astype(np.float32)
            f = torch.arange(ec, dtype=torch.float32).reshape(1, ec)
            out_ref, lg_ref = O.pool_mesh(ref, f, T, norms=norms)
            lg, rounds, examined, mask, groups = _reservation_pool(par, norms, T, window)
            np.testing.assert_array_equal(ref.gemm_edges, par.gemm_edges)
            np.testing.assert_array_equal(ref.sides, par.sides)
            np.testing.assert_array_equal(ref.edges, par.edges)
            np.testing.assert_array_equal(ref.v_mask, par.v_mask)
            assert ref.ve == par.ve and ref.edges_count == par.edges_count
            np.testing.assert_array_equal(lg_ref['mask'], mask)
            rows = groups.clamped_rows(mask)
            assert len(rows) == len(lg_ref['rows']) and all(np.array_equal(a, b) for a, b in zip(rows, lg_ref['rows']))
            np.testing.assert_array_equal(groups.occurrences(), lg_ref['occurrences'])
            assert sorted(lg['pops']) == sorted(p for p, o in zip(lg_ref['pops'], lg_ref['outcomes']) if o != O.SKIP_MASKED)
            tot_rounds += rounds
            tot_exam += examined
    print('\n  window %d: %d candidates examined in %d rounds (%.2f per round)' % (window, tot_exam, tot_rounds, tot_exam / tot_rounds))
