"""Executable specification of the collapse kernel's schedule (csrc/meshpool.cu, phase D), checked on the CPU oracle.

The kernel does not pop one edge at a time.  It works in ROUNDS over a window of the next W unmasked candidates of the sorted
order (one per warp):
  inspect   read-only on the committed state: has_boundaries, the __get_invalids test of both sides (without its side
            effects), __is_one_ring_valid  ->  verdict REJ_BOUNDARY / REJ_ONE_RING / COLLAPSE / ALONE (needs rewiring);
            the rows read: edge rows {e, its four neighbours}, vertex rows {v0, v1, every endpoint of every entry of the two
            vertex lists}
  reserve   a COLLAPSE candidate tags the rows it will read again and write: edge rows {e, key_a, key_b and key_b's two
            far-face neighbours, per side}, vertex rows {v0, v1, the endpoints of both key_b}; the earliest candidate wins
  commit    the longest prefix of the window whose members touch (read or write) no row tagged by an earlier member, and
            which the edge count allows (mesh_pool.py:49), is committed in parallel; an ALONE candidate commits only at the
            head of the window, through the complete sequential __pool_edge
Every committed member has read exactly what it would have read in the strict order, so the result must be IDENTICAL to the
reference's one-at-a-time loop: pops, outcomes, union log (in order), topology, vertex lists, masks, groups.  This file
re-states that schedule on the oracle's data structures and asserts the identity; tests/test_pool_gpu.py asserts it for the
kernel itself.  (tools/pool_wavefront_study.py holds the dependency study that led to this design.)"""
import numpy as np
import pytest
import torch

from meshcnn_b200 import synth
from oracle import meshcnn_oracle as O
from tests.helpers import prepared

BOUNDARY, ONE_RING, COLLAPSE, ALONE = 1, 2, 3, 4


def _inspect(run, e):
    """-> (verdict, edge rows read, vertex rows read, edge rows written, vertex rows written)"""
    m = run.m
    g, s, ed = m.gemm_edges, m.sides, m.edges
    row = [int(x) for x in g[e]]
    if any(x < 0 for x in row):
        return BOUNDARY, {e}, set(), set(), set()
    re_ = {e} | set(row)
    if any(int(x) < 0 for r in row for x in g[r]):
        return BOUNDARY, re_, set(), set(), set()
    alone = False
    okb = []
    for side in (0, 2):
        ka, kb, sa, sb, osa, osb, oka, okb_ = run.face_info(e, side)
        alone |= any(a == b for a in oka for b in okb_)
        okb.append(okb_)
    v0, v1 = int(ed[e, 0]), int(ed[e, 1])
    alone |= v0 == v1
    if alone:
        return ALONE, re_, set(), set(), set()
    rv = {v0, v1}
    for x in list(m.ve[v0]) + list(m.ve[v1]):
        rv.update(int(t) for t in ed[x])
    if not run.one_ring_valid(e):
        return ONE_RING, re_, rv, set(), set()
    we = {e, row[0], row[1], okb[0][0], okb[0][1], row[2], row[3], okb[1][0], okb[1][1]}
    wv = {v0, v1} | {int(t) for t in ed[row[1]]} | {int(t) for t in ed[row[3]]}
    return COLLAPSE, re_, rv, we, wv


def _rounds_pool(mesh, norms, target, W):
    ec0 = mesh.edges_count
    order = [int(e) for e in O.collapse_order(norms, ec0)]
    groups = O.SparseGroups(ec0)
    log = {'pops': [], 'outcomes': [], 'unions': [], 'triplets': []}
    run = O._PoolRun(mesh, target, groups, log)
    cursor, rounds, committed = 0, 0, 0
    while mesh.edges_count > target:
        win = [p for p in range(cursor, ec0) if run.mask[order[p]]][:W]
        if not win:
            log['pops'] += order[cursor:]
            log['outcomes'] += [O.SKIP_MASKED] * (ec0 - cursor)
            raise O.PoolExhausted('pop from an empty queue')
        insp = [_inspect(run, order[p]) for p in win]
        own_e, own_v = {}, {}
        for j, (v, re_, rv, we, wv) in enumerate(insp):
            for x in we:
                own_e.setdefault(x, j)
            for x in wv:
                own_v.setdefault(x, j)
        k, ec_run = 0, mesh.edges_count
        for j, (v, re_, rv, we, wv) in enumerate(insp):
            if ec_run <= target:
                break
            if v != ALONE and (any(own_e.get(x, j) < j for x in re_ | we) or any(own_v.get(x, j) < j for x in rv | wv)):
                break
            if v == ALONE:
                if j == 0:
                    k = 1
                break
            k = j + 1
            if v == COLLAPSE:
                ec_run -= 3
        assert k >= 1
        # commit members 0 .. k-1; the log is written in candidate order == the strict order
        outcomes = {}
        for j in range(k):
            e, v = order[win[j]], insp[j][0]
            if v == ALONE:
                outcomes[win[j]] = run.pool_edge(e)
            elif v == COLLAPSE:
                run.pool_side(e, 0)
                run.pool_side(e, 2)
                mesh.merge_vertices(e)
                run.mask[e] = False
                mesh.edges_count -= 1
                outcomes[win[j]] = O.COLLAPSED
            else:
                outcomes[win[j]] = O.REJ_BOUNDARY if v == BOUNDARY else O.REJ_ONE_RING
        new_cursor = win[k - 1] + 1
        for p in range(cursor, new_cursor):
            log['pops'].append(order[p])
            log['outcomes'].append(outcomes.get(p, O.SKIP_MASKED))
        cursor = new_cursor
        rounds += 1
        committed += k
    mesh.clean(run.mask, groups)
    return log, run.mask, groups, rounds, committed


CASES = [('ico', dict(nu=5), [600, 450, 300, 210, 100, 40]), ('torus', dict(m=10, n=25), [600, 450, 300, 180, 90]),
         ('grid', dict(m=10, n=12), [330, 280, 230])]


@pytest.mark.parametrize('W', [2, 8, 32])
@pytest.mark.parametrize('kind,kw,chain', CASES)
def test_round_schedule_reproduces_the_strict_order(kind, kw, chain, W):
    stats = [0, 0]
    for s in range(3):
        vs, faces = synth.make(kind, s, **kw)
        data = prepared(vs, faces)
        ref, par = O.OracleMesh(data), O.OracleMesh(data)
        rs = np.random.RandomState(300 + s)
        for T in chain:
            ec = ref.edges_count
            # smooth-ish priorities on odd levels (neighbouring ids get similar norms): more conflicts inside a window
            norms = rs.rand(ec).astype(np.float32)
            if len(stats) and (T // 10) % 2:
                norms = (np.arange(ec) / ec + 0.05 * rs.rand(ec)).astype(np.float32)
            f = torch.arange(ec, dtype=torch.float32).reshape(1, ec)
            try:
                _, lg_ref = O.pool_mesh(ref, f, T, norms=norms)
            except O.PoolExhausted:
                with pytest.raises(O.PoolExhausted):
                    _rounds_pool(par, norms, T, W)
                break
            lg, mask, groups, rounds, committed = _rounds_pool(par, norms, T, W)
            assert lg['pops'] == lg_ref['pops']
            assert lg['outcomes'] == lg_ref['outcomes']
            assert lg['unions'] == lg_ref['unions']                      # the union log, in order
            assert lg['triplets'] == lg_ref['triplets']
            np.testing.assert_array_equal(ref.gemm_edges, par.gemm_edges)
            np.testing.assert_array_equal(ref.sides, par.sides)
            np.testing.assert_array_equal(ref.edges, par.edges)
            np.testing.assert_array_equal(ref.v_mask, par.v_mask)
            np.testing.assert_array_equal(ref.vs, par.vs)
            assert ref.ve == par.ve and ref.edges_count == par.edges_count
            np.testing.assert_array_equal(lg_ref['mask'], mask)
            rows = groups.clamped_rows(mask)
            assert len(rows) == len(lg_ref['rows']) and all(np.array_equal(a, b) for a, b in zip(rows, lg_ref['rows']))
            np.testing.assert_array_equal(groups.occurrences(), lg_ref['occurrences'])
            stats[0] += rounds
            stats[1] += committed
    print('\n  %s W=%d: %d candidates committed in %d rounds (%.2f per round)' % (kind, W, stats[1], stats[0], stats[1] / max(1, stats[0])))
