"""Mesh (host object, drop-in for models/layers/mesh.py) and MeshBatch (device-resident topology).

Reference behaviour mirrored here (file:line into the reference):
  * ``Mesh(file, opt, hold_history, export_folder)`` with attributes ``vs, v_mask, filename, features,
    edge_areas, edge_lengths, edges, gemm_edges, sides, ve, edges_count, pool_count, export_folder,
    history_data``                                                                     mesh.py:12-21
  * meshes are built in DataLoader workers and pickled (data/base_dataset.py:55-63), so a Mesh is
    plain host data at birth; the device copy is made lazily, once per batch, by ``MeshBatch.of``
  * a Mesh is single-use per forward: MeshUnpool restores only gemm_edges / edges_count
                                                                                       mesh.py:194-198

B200-first design: the topology of the whole batch lives in HBM as flat arrays (one upload per batch, one
blob); MeshPool rewrites it on the device and the host attributes become lazy mirrors that are downloaded only
when host code actually reads them (``Mesh.gemm_edges`` etc. are properties).
"""
import os

import numpy as np
import torch

from . import mesh_prepare
from . import _lib
from . import profiler


class PoolError(RuntimeError):
    pass


# status word -> exception the reference would have raised
_STATUS_EXC = {
    1: (IndexError, 'pop from an empty heap: the edge queue ran dry before the pool target was reached (mesh_pool.py:49-50)'),
    2: (AssertionError, 'assert len(shared_items) == 2 (mesh_pool.py:123)'),
    3: (AssertionError, 'assert len(vertex) == 1 (mesh_pool.py:181)'),
    4: (ValueError, 'list.remove(x): x not in list (mesh.py:48)'),
    5: (PoolError, 'MeshUnion node pool exhausted (implementation limit of the collapse kernel)'),
    6: (PoolError, 'degenerate edge (identical endpoints) reached merge_vertices'),
    7: (PoolError, 'groups CSR capacity exceeded'),
}


def _align(n, a=16):
    return (n + a - 1) // a * a


_PIN_POOL = {}          # size -> list of [pinned tensor, cuda event]; cudaHostAlloc per batch would cost milliseconds


def _staging_buffer(nbytes, device):
    """A pinned host buffer for the topology upload, recycled across batches (two per size, so packing batch k+1 can
    overlap the copy of batch k)."""
    if torch.device(device).type != 'cuda':
        return torch.empty(nbytes, dtype=torch.uint8), None
    slots = _PIN_POOL.setdefault(nbytes, [])
    for slot in slots:
        if slot[1].query():
            return slot[0], slot
    if len(slots) >= 2:
        slots[0][1].synchronize()
        return slots[0][0], slots[0]
    slot = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), torch.cuda.Event()]
    slots.append(slot)
    return slot[0], slot


# Collapse-kernel failures (the exceptions the reference raises inside MeshPool) must reach a plain training loop that never
# reads a mirrored Mesh attribute, without a device synchronisation per step: every pool() call copies its per-mesh status
# words into pinned host memory asynchronously; the words are looked at as soon as the copy has completed -- at the next
# MeshBatch construction / load / pool call, i.e. at the latest in the following step -- and raise there.
_STATUS_LEVELS = 16
_STATUS_FREE = {}        # B -> recycled pinned [_STATUS_LEVELS, B] int32 buffers
_STATUS_PENDING = []     # [event, pinned buffer, number of levels written, filenames]


def _raise_status(words, names):
    for lvl in range(words.shape[0]):
        bad = np.nonzero(words[lvl])[0]
        if len(bad):
            b = int(bad[0])
            exc, msg = _STATUS_EXC.get(int(words[lvl, b]), (PoolError, 'status %d' % int(words[lvl, b])))
            raise exc('MeshPool (pool call %d of the forward), mesh %d of the batch (%s): %s' % (lvl, b, names[b], msg))


def poll_pool_status(block=False):
    """Look at the status words of finished pool calls (all of them with block=True); raises what the reference would have
    raised inside MeshPool.forward for the first failed mesh."""
    i = 0
    while i < len(_STATUS_PENDING):
        ent = _STATUS_PENDING[i]
        ev, buf, n, names = ent
        if block:
            ev.synchronize()
        if not ev.query():
            i += 1
            continue
        _STATUS_PENDING.pop(i)
        words = buf[:n].numpy().copy()
        ent[2] = 0
        _STATUS_FREE.setdefault(buf.shape[1], []).append(buf)
        _raise_status(words, names)


class Level:
    """One resolution of the batch: one-ring tables + live edge counts, all on the device."""
    __slots__ = ('E', 'gemm', 'sides', 'ec', 'padded')

    def __init__(self, E, gemm, sides, ec):
        self.E, self.gemm, self.sides, self.ec = E, gemm, sides, ec
        self.padded = {}

    def for_width(self, E):
        """Tables for an activation tensor whose edge dimension is E >= self.E (the reference pads per call,
        mesh_conv.py:82)."""
        if E == self.E:
            return self
        if E < self.E:
            raise ValueError('activation has %d edge slots but the mesh level has %d' % (E, self.E))
        lv = self.padded.get(E)
        if lv is None:
            B = self.gemm.shape[0]
            g = torch.zeros(B, E, 4, dtype=torch.int32, device=self.gemm.device)
            s = torch.zeros(B, E, 4, dtype=torch.int8, device=self.gemm.device)
            g[:, :self.E] = self.gemm
            s[:, :self.E] = self.sides
            lv = Level(E, g, s, self.ec)
            self.padded[E] = lv
        return lv


class PoolRecord:
    """What one MeshPool call leaves behind for MeshUnpool / backward: MeshUnion rows as CSR + CSC and the
    un-clamped occurrences (mesh.py:182-192)."""
    __slots__ = ('E_in', 'T', 'nnz_cap', 'rowptr', 'cols', 'colptr', 'rows', 'occ', 'status', 'level_in', 'level_out',
                 'norms', 'logs', 'edge_mask', 'clocks')


class MeshBatch:
    """Device state of a batch of meshes."""

    def __init__(self, meshes, E, device, with_vs=None, V_cap=0, ve_cap=0, staging=None):
        """meshes: B Mesh objects; E: edge slots of the level-0 activations.  V_cap / ve_cap reserve room for later
        `load()` calls with other meshes (CUDA-graph replay); staging: a caller-owned pinned uint8 tensor to pack into
        (so that the host->device copy can itself be a node of a captured graph)."""
        meshes = list(meshes)
        self.device = torch.device(device)
        if _STATUS_PENDING:
            poll_pool_status()                                   # failures of earlier steps surface here, one step late at most
        B = len(meshes)
        hs = [m._host_state() for m in meshes]
        V = max(max(len(h['vs']) for h in hs), int(V_cap))
        nve = max(int(len(h['ve_flat'])) for h in hs)
        ve_cap = _align(max(2 * E, nve, int(ve_cap)), 8)
        if with_vs is None:
            with_vs = any(bool(m.export_folder) for m in meshes)
        self.B, self.E0, self.V, self.ve_cap, self.with_vs = B, E, V, ve_cap, with_vs
        sizes = [('gemm', B * E * 4 * 4), ('sides', B * E * 4), ('ec', B * 4), ('edges', B * E * 2 * 4),
                 ('ve_ptr', B * (V + 1) * 4), ('ve_idx', B * ve_cap * 4), ('v_mask', B * V),
                 ('vs', B * V * 3 * 8 if with_vs else 0)]
        offs, o = {}, 0
        for k, s in sizes:
            offs[k] = (o, s)
            o = _align(o + s)
        self._offs, self.nbytes = offs, o
        self._staging = staging
        if staging is not None and (staging.numel() < o or not staging.is_pinned()):
            raise ValueError('staging buffer must be pinned and hold %d bytes' % o)
        self.blob = torch.empty(o, dtype=torch.uint8, device=self.device)
        self.meshes = []
        self.version = 0
        self._upload(meshes, hs)

        def dview(k, dtype, shape):
            a, s = offs[k]
            return self.blob[a:a + s].view(dtype).view(shape)

        self.edges = dview('edges', torch.int32, (B, E, 2))
        self.ve_ptr = dview('ve_ptr', torch.int32, (B, V + 1))
        self.ve_idx = dview('ve_idx', torch.int32, (B, ve_cap))
        self.v_mask = dview('v_mask', torch.uint8, (B, V))
        self.vs = dview('vs', torch.float64, (B, V, 3)) if with_vs else None
        self.level0 = Level(E, dview('gemm', torch.int32, (B, E, 4)), dview('sides', torch.int8, (B, E, 4)),
                            dview('ec', torch.int32, (B,)))
        self.levels = [self.level0]
        self.cur = 0
        self.stack = []            # PoolRecords not yet undone by MeshUnpool
        self.records = []          # every PoolRecord of this forward, in order
        self._pristine = None
        self.keep_logs = False
        self.keep_clocks = False
        self.force_wide = os.environ.get('MESHCNN_B200_POOL_WIDE', '0') == '1'    # tests: large-mesh variant on small meshes
        self._pool_ws = None
        self._status_ent = None    # entry of _STATUS_PENDING (eager) or [None, pinned buffer, n, names] (captured step)

    def pack(self, meshes, hs=None):
        """Pack the topology of `meshes` into the pinned staging buffer (one blob); returns it.  No device work."""
        meshes = list(meshes)
        B, E, V, ve_cap, with_vs, offs = self.B, self.E0, self.V, self.ve_cap, self.with_vs, self._offs
        if len(meshes) != B:
            raise ValueError('this MeshBatch holds %d meshes, got %d' % (B, len(meshes)))
        if hs is None:
            hs = [m._host_state() for m in meshes]
        ecs = [int(h['edges_count']) for h in hs]
        if max(ecs) > E:
            raise ValueError('mesh with %d edges does not fit an activation of width %d' % (max(ecs), E))
        if max(len(h['vs']) for h in hs) > V or max(len(h['ve_flat']) for h in hs) > ve_cap:
            raise ValueError('mesh exceeds the vertex / list capacity this MeshBatch was built with (V_cap, ve_cap)')
        if self._staging is not None:
            host, self._pin_slot = self._staging, None
        else:
            host, self._pin_slot = _staging_buffer(self.nbytes, self.device)
        hb = host.numpy()[:self.nbytes]
        hb[:] = 0

        def view(k, dtype, shape):
            a, s = offs[k]
            return hb[a:a + s].view(dtype).reshape(shape)

        gemm = view('gemm', np.int32, (B, E, 4))
        sides = view('sides', np.int8, (B, E, 4))
        ec = view('ec', np.int32, (B,))
        edges = view('edges', np.int32, (B, E, 2))
        ve_ptr = view('ve_ptr', np.int32, (B, V + 1))
        ve_idx = view('ve_idx', np.int32, (B, ve_cap))
        v_mask = view('v_mask', np.uint8, (B, V))
        vs = view('vs', np.float64, (B, V, 3)) if with_vs else None
        for b, h in enumerate(hs):
            n = ecs[b]
            gemm[b, :n] = h['gemm_edges'][:n]
            sides[b, :n] = h['sides'][:n]
            ec[b] = n
            edges[b, :n] = h['edges'][:n]
            off = h['ve_off']
            ve_ptr[b, :len(off)] = off
            ve_ptr[b, len(off):] = off[-1]
            ve_idx[b, :len(h['ve_flat'])] = h['ve_flat']
            v_mask[b, :len(h['v_mask'])] = h['v_mask']
            if with_vs:
                vs[b, :len(h['vs'])] = h['vs']
        self._host_blob = host
        for m in self.meshes:                                  # the previous occupants keep their last host mirror
            if getattr(m, '_batch', None) is self:
                m._batch = None
        self.meshes = meshes
        for i, m in enumerate(meshes):
            m._attach(self, i)
        return host

    def _upload(self, meshes, hs=None):
        host = self.pack(meshes, hs)
        self.blob.copy_(host[:self.nbytes], non_blocking=True)
        if self._pin_slot is not None:
            self._pin_slot[1].record()                 # the staging buffer may be reused once this copy has completed

    def load(self, meshes):
        """Re-use the device buffers of this batch for another list of B meshes: one packed upload, state back at
        level 0.  Everything that was computed from the previous occupants is dropped."""
        self._upload(meshes)
        self.rewind()

    def rewind(self):
        """Host-side bookkeeping back to level 0 (no device work)."""
        if _STATUS_PENDING and not torch.cuda.is_current_stream_capturing():
            poll_pool_status()
        self.levels = [self.level0]
        self.level0.padded.clear()                               # tables widened for another batch / another replay are stale
        self.cur = 0
        self.stack, self.records = [], []
        self.version += 1
        for m in self.meshes:
            m.pool_count = 0

    # -- lookup ------------------------------------------------------------------------------------------
    @staticmethod
    def of(meshes, E, device):
        """The MeshBatch the given meshes are attached to (created on first use: one upload per batch)."""
        first = meshes[0]
        bt = getattr(first, '_batch', None)
        if bt is not None and bt.device == torch.device(device) and len(bt.meshes) == len(meshes) \
                and all(getattr(m, '_batch', None) is bt and m._bi == i for i, m in enumerate(meshes)):
            return bt
        return MeshBatch(meshes, E, device)

    def level(self, E=None):
        lv = self.levels[self.cur]
        return lv if E is None else lv.for_width(E)

    # -- single-use handling for benchmarks ----------------------------------------------------------------
    def snapshot(self):
        """Keep a device copy of the carried state so the batch can be re-used (benchmarks, CUDA graphs)."""
        a = self._offs['edges'][0]
        self._pristine = self.blob[a:].clone()

    def reset(self):
        """Restore the state saved by snapshot(): level 0, original edges / ve / v_mask / vs."""
        if self._pristine is None:
            raise RuntimeError('MeshBatch.reset() without snapshot()')
        a = self._offs['edges'][0]
        self.blob[a:].copy_(self._pristine)
        self.rewind()

    # -- pool / unpool bookkeeping -------------------------------------------------------------------------
    def pool(self, norms, T):
        """Run the collapse kernel for the current level; returns the PoolRecord (topology advanced)."""
        L = _lib.lib()
        lv = self.levels[self.cur]
        B, E_in, dev = self.B, lv.E, self.device
        need = L.mcnn_pool_smem_bytes(E_in, self.V, self.ve_cap, 1 if self.with_vs else 0)
        # meshes whose collapse state does not fit on chip run the same kernel on a global-memory workspace (32-bit ids)
        wide = self.force_wide or need > L.mcnn_pool_smem_limit() or E_in > 10900 or self.V > 32767 or self.ve_cap > 65535
        ws_stride = 0
        if wide:
            ws_stride = L.mcnn_pool_workspace_bytes(E_in, self.V, self.ve_cap, 1 if self.with_vs else 0)
            if self._pool_ws is None or self._pool_ws.numel() < B * ws_stride:
                self._pool_ws = torch.empty(B * ws_stride, dtype=torch.uint8, device=dev)
        nnz_cap = 3 * E_in
        ints = torch.empty(B * (T * 4 + 1 + (T + 1) + (E_in + 1) + 2 * nnz_cap + 1), dtype=torch.int32, device=dev)
        o = 0

        def take(n, shape):
            nonlocal o
            t = ints[o:o + n].view(shape)
            o += n
            return t

        gemm_out = take(B * T * 4, (B, T, 4))
        ec_out = take(B, (B,))
        rowptr = take(B * (T + 1), (B, T + 1))
        colptr = take(B * (E_in + 1), (B, E_in + 1))
        cols = take(B * nnz_cap, (B, nnz_cap))
        rows = take(B * nnz_cap, (B, nnz_cap))
        status = take(B, (B,))
        sides_out = torch.empty(B, T, 4, dtype=torch.int8, device=dev)
        occ = torch.empty(B, E_in, dtype=torch.float32, device=dev)
        rec = PoolRecord()
        rec.E_in, rec.T, rec.nnz_cap = E_in, T, nnz_cap
        rec.rowptr, rec.cols, rec.colptr, rec.rows, rec.occ, rec.status = rowptr, cols, colptr, rows, occ, status
        rec.norms = norms
        rec.logs = None
        rec.edge_mask = None
        a = _lib.PoolArgs()
        a.B, a.E_in, a.T, a.V, a.ve_cap, a.nnz_cap = B, E_in, T, self.V, self.ve_cap, nnz_cap
        a.log_cap = 0
        a.gemm_in, a.sides_in, a.ec_in, a.norms = lv.gemm.data_ptr(), lv.sides.data_ptr(), lv.ec.data_ptr(), norms.data_ptr()
        a.edges, a.edges_stride = self.edges.data_ptr(), self.E0
        a.ve_ptr, a.ve_idx, a.v_mask = self.ve_ptr.data_ptr(), self.ve_idx.data_ptr(), self.v_mask.data_ptr()
        a.vs = self.vs.data_ptr() if self.with_vs else None
        a.gemm_out, a.sides_out, a.ec_out = gemm_out.data_ptr(), sides_out.data_ptr(), ec_out.data_ptr()
        a.grp_rowptr, a.grp_cols, a.grp_colptr, a.grp_rows = rowptr.data_ptr(), cols.data_ptr(), colptr.data_ptr(), rows.data_ptr()
        a.occ, a.status = occ.data_ptr(), status.data_ptr()
        if wide:
            a.workspace, a.workspace_stride, a.force_workspace = self._pool_ws.data_ptr(), ws_stride, 1
        exporting = [i for i, m in enumerate(self.meshes) if m.export_folder]
        if (exporting or any(m.history_data is not None for m in self.meshes)) and not self.keep_logs:
            rec.edge_mask = torch.empty(B, E_in, dtype=torch.int32, device=dev)      # the kernel writes every entry
            a.edge_mask_out = rec.edge_mask.data_ptr()
        if self.keep_logs:
            cap = E_in
            lg = {'npops': torch.zeros(B, dtype=torch.int32, device=dev),
                  'pops': torch.zeros(B, cap, dtype=torch.int32, device=dev),
                  'outcome': torch.zeros(B, cap, dtype=torch.int8, device=dev),
                  'nunions': torch.zeros(B, dtype=torch.int32, device=dev),
                  'unions': torch.zeros(B, cap * 8, 2, dtype=torch.int32, device=dev)}
            rec.edge_mask = torch.zeros(B, E_in, dtype=torch.int32, device=dev)
            a.log_cap = cap
            a.log_npops, a.log_pops, a.log_outcome = lg['npops'].data_ptr(), lg['pops'].data_ptr(), lg['outcome'].data_ptr()
            a.log_nunions, a.log_unions = lg['nunions'].data_ptr(), lg['unions'].data_ptr()
            a.edge_mask_out = rec.edge_mask.data_ptr()
            rec.logs = lg
        rec.clocks = None
        if self.keep_clocks:
            rec.clocks = torch.zeros(B, 24, dtype=torch.int64, device=dev)
            a.dbg_clocks = rec.clocks.data_ptr()
        with profiler.span('pool_collapse', B=B, E_in=E_in, T=T):
            _lib.check(L.mcnn_pool_collapse(a, _lib.current_stream()), 'mcnn_pool_collapse')
        self._publish_status(status)
        new = Level(T, gemm_out, sides_out, ec_out)
        rec.level_in = self.cur
        self.levels = self.levels[:self.cur + 1] + [new]
        self.cur += 1
        rec.level_out = self.cur
        self.stack.append(rec)
        self.records.append(rec)
        self.version += 1
        for m in self.meshes:
            m.pool_count += 1
        for i in exporting:                              # Mesh.clean ends with export() (mesh.py:71): test / visualisation
            m = self.meshes[i]                           # runs only -- this synchronises and downloads the level
            if m.history_data is not None and 'edges_mask' in m.history_data:
                alive = m.history_data['edges_mask'][-1]
                keep = rec.edge_mask[i, :int(alive.sum())].cpu().numpy().astype(bool)
                idx = np.nonzero(alive)[0]
                alive[idx[~keep]] = False
                m.history_data['edges_mask'].append(alive.copy())
            m.export()
        return rec

    def _publish_status(self, status):
        """Asynchronous copy of one pool call's status words into pinned memory (see poll_pool_status)."""
        if self.device.type != 'cuda':
            return
        k = len(self.records)
        if k >= _STATUS_LEVELS:
            return
        capturing = torch.cuda.is_current_stream_capturing()
        ent = self._status_ent
        if capturing:
            # the copy becomes a node of the graph: a buffer of its own, never recycled, read by check_captured_status()
            if ent is None or ent[0] is not None:
                ent = self._status_ent = [None, torch.zeros(_STATUS_LEVELS, self.B, dtype=torch.int32).pin_memory(), 0, None]
        elif ent is None or ent[0] is None or not any(ent is e for e in _STATUS_PENDING):
            free = _STATUS_FREE.get(self.B)
            buf = free.pop() if free else torch.zeros(_STATUS_LEVELS, self.B, dtype=torch.int32).pin_memory()
            ent = self._status_ent = [torch.cuda.Event(), buf, 0, None]
            _STATUS_PENDING.append(ent)
        ent[1][k].copy_(status, non_blocking=True)
        ent[2] = max(ent[2], k + 1)
        ent[3] = [m.filename for m in self.meshes]
        if not capturing:
            ent[0].record()

    def check_captured_status(self):
        """For a step replayed from a CUDA graph: the status words of the LAST COMPLETED replay sit in pinned memory (the
        copies are graph nodes); the caller guarantees that replay has finished (GraphedTrainStep waits on its event)."""
        ent = self._status_ent
        if ent is not None and ent[2]:
            _raise_status(ent[1][:ent[2]].numpy().copy(), ent[3])

    def unroll(self):
        """Mesh.get_groups / get_occurrences / unroll_gemm (mesh.py:176-198): pop one level."""
        if not self.stack:
            raise IndexError('pop from empty list')          # what history_data['groups'].pop() raises
        rec = self.stack.pop()
        self.cur = rec.level_in
        self.version += 1
        return rec

    def check(self):
        """Synchronise and raise the exception the reference would have raised inside MeshPool (deferred)."""
        ent = self._status_ent
        if ent is not None:                                  # this is the full check: nothing left for the deferred one
            _STATUS_PENDING[:] = [e for e in _STATUS_PENDING if e is not ent]
            self._status_ent = None
        for rec in self.records:
            st = rec.status.cpu().numpy()
            bad = np.nonzero(st)[0]
            if len(bad):
                exc, msg = _STATUS_EXC.get(int(st[bad[0]]), (PoolError, 'status %d' % int(st[bad[0]])))
                raise exc('MeshPool, mesh %d of the batch (%s): %s' % (int(bad[0]), self.meshes[int(bad[0])].filename, msg))

    # -- host mirrors ----------------------------------------------------------------------------------------
    def download(self, i):
        """Current device state of mesh i as the reference's host attributes."""
        self.check()
        lv = self.levels[self.cur]
        n = int(lv.ec[i].item())
        out = {'edges_count': n, 'gemm_edges': lv.gemm[i, :n].cpu().numpy().astype(np.int64)}
        # sides / edges / ve / v_mask / vs follow the most recent pool (they are not restored by unroll_gemm)
        deep = self.levels[-1]
        nd = int(deep.ec[i].item())
        out['sides'] = deep.sides[i, :nd].cpu().numpy().astype(np.int64)
        out['edges'] = self.edges[i, :nd].cpu().numpy().astype(np.int32)
        nv = len(self.meshes[i]._h['vs'])
        ptr = self.ve_ptr[i, :nv + 1].cpu().numpy()
        out['ve_off'] = ptr.astype(np.int32)
        out['ve_flat'] = self.ve_idx[i, :int(ptr[-1])].cpu().numpy().astype(np.int32)
        out['v_mask'] = self.v_mask[i, :nv].cpu().numpy().astype(bool)
        if self.with_vs:
            out['vs'] = self.vs[i, :nv].cpu().numpy()
        return out


def _flatten_ve(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    if len(lists):
        off[1:] = np.cumsum([len(l) for l in lists])
    flat = np.fromiter((e for l in lists for e in l), dtype=np.int32, count=int(off[-1]))
    return flat, off


class Mesh:
    """Drop-in for models/layers/mesh.py:Mesh (host side)."""

    _MIRRORED = ('gemm_edges', 'sides', 'edges', 'v_mask', 'vs', 'edges_count')

    def __init__(self, file=None, opt=None, hold_history=False, export_folder='', data=None):
        if data is None:
            if file is None:
                raise ValueError('Mesh needs an .obj file or prepared data')
            data = mesh_prepare.fill(file, opt)              # mesh.py:16 fill_mesh: npz cache + augmentation per `opt`
        self._h = {'vs': np.array(data.vs, dtype=np.float64), 'v_mask': np.array(data.v_mask, dtype=bool),
                   'edges': np.array(data.edges, dtype=np.int32), 'gemm_edges': np.array(data.gemm_edges, dtype=np.int64),
                   'sides': np.array(data.sides, dtype=np.int64), 'edges_count': int(data.edges_count)}
        self._h['ve_flat'], self._h['ve_off'] = _flatten_ve(data.ve)
        self.filename = data.filename
        self.features = data.features
        self.edge_areas = data.edge_areas
        self.edge_lengths = data.edge_lengths
        self.pool_count = 0
        self.export_folder = export_folder
        self.history_data = None
        self._batch, self._bi, self._seen = None, -1, -1
        self._hist_seen = 0
        if hold_history:
            self.init_history()
        self.export()                                    # mesh.py:21 (a no-op without an export folder)

    # ---- device link -------------------------------------------------------------------------------------
    def _attach(self, batch, i):
        self._batch, self._bi, self._seen = batch, i, batch.version
        self._hist_seen = 0

    def _sync(self):
        bt = self._batch
        if bt is not None and self._seen != bt.version:
            self._h.update(bt.download(self._bi))
            self._h.pop('ve_lists', None)
            self._seen = bt.version

    def _host_state(self):
        self._sync()
        lists = self._h.get('ve_lists')
        if lists is not None:                            # host code may have edited the lists in place
            self._h['ve_flat'], self._h['ve_off'] = _flatten_ve(lists)
        return self._h

    def __getstate__(self):
        self._host_state()
        self._pull_history()
        d = dict(self.__dict__)
        d['_batch'], d['_bi'], d['_seen'] = None, -1, -1
        d['_hist_seen'] = 0
        return d

    def _own(self):
        """Host code is about to mutate this mesh (the methods of mesh.py:26-71): fetch the current device state, mirror the
        history the device holds, and detach -- from here on the host copy is the truth (a later forward re-uploads it)."""
        if self._batch is None:
            return
        self._host_state()
        self._pull_history()
        self._batch = None

    # ---- reference API -------------------------------------------------------------------------------------
    @property
    def ve(self):
        """vertex -> incident edge ids as a list of lists (mesh_prepare.py:142-143).  Stored flat (CSR) so that a batch can
        be packed for upload without touching Python lists; the list form is made on first read and then kept, so in-place
        edits (``mesh.ve[v].remove(e)``) stick like on the reference's attribute."""
        self._sync()
        lists = self._h.get('ve_lists')
        if lists is None:
            flat, off = self._h['ve_flat'], self._h['ve_off']
            lists = self._h['ve_lists'] = [flat[off[v]:off[v + 1]].tolist() for v in range(len(off) - 1)]
        return lists

    @ve.setter
    def ve(self, lists):
        self._sync()
        self._h['ve_lists'] = [list(l) for l in lists]
        self._h['ve_flat'], self._h['ve_off'] = _flatten_ve(lists)

    def extract_features(self):                      # mesh.py:23-24
        return self.features

    def get_edge_areas(self):                        # mesh.py:200-201
        return self.edge_areas

    # ---- host-side mutation (mesh.py:26-71).  MeshPool does all of this on the device; these are the reference's methods
    # for host code that edits a mesh itself.  They work on the host mirror and detach the mesh from its device batch. ----
    def merge_vertices(self, edge_id):               # mesh.py:26-37
        self._own()
        self.remove_edge(edge_id)
        h = self._h
        v0, v1 = int(h['edges'][edge_id][0]), int(h['edges'][edge_id][1])
        h['vs'][v0] += h['vs'][v1]
        h['vs'][v0] /= 2
        h['v_mask'][v1] = False
        hit = h['edges'] == v1                       # ALL rows, dead and stale ones included
        self.ve[v0].extend(self.ve[v1])
        h['edges'][hit] = v0

    def remove_vertex(self, v):                      # mesh.py:39-40
        self._own()
        self._h['v_mask'][v] = False

    def remove_edge(self, edge_id):                  # mesh.py:42-48 (list.remove: first occurrence, ValueError if absent)
        self._own()
        for v in self._h['edges'][edge_id]:
            self.ve[int(v)].remove(edge_id)

    def clean(self, edges_mask, groups):             # mesh.py:50-71
        self._own()
        h = self._h
        keep = np.asarray(edges_mask).astype(bool)
        h['gemm_edges'] = h['gemm_edges'][:len(keep)][keep]
        h['edges'] = h['edges'][:len(keep)][keep]
        h['sides'] = h['sides'][:len(keep)][keep]
        renum = np.zeros(len(keep) + 1, dtype=np.int32)          # dead -> 0; the extra last slot serves id -1
        renum[-1] = -1
        renum[:-1][keep] = np.arange(int(keep.sum()))
        h['gemm_edges'][:, :] = renum[h['gemm_edges']]
        self.ve = [[int(renum[e]) for e in l] for l in self.ve]
        self.__clean_history(groups, torch.from_numpy(keep.copy()))
        self.pool_count += 1
        self.export()

    # ---- history (mesh.py:151-198) ---------------------------------------------------------------------------
    def init_history(self):                          # mesh.py:151-162
        n = self._h['edges_count']
        self.history_data = {'groups': [], 'gemm_edges': [self._h['gemm_edges'].copy()], 'occurrences': [],
                             'old2current': np.arange(n, dtype=np.int32), 'current2old': np.arange(n, dtype=np.int32),
                             # survivors as a mask over the ORIGINAL edge ids; one entry per exported level (export_segments)
                             'edges_mask': [np.ones(n, dtype=bool)],
                             'edges_count': [n]}
        self._hist_seen = 0                          # pool records of the device batch already mirrored into the lists
        if self.export_folder:
            from .mesh_union import MeshUnion
            self.history_data['collapses'] = MeshUnion(n)

    def union_groups(self, source, target):          # mesh.py:164-167
        if self.export_folder and self.history_data:
            c2o = self.history_data['current2old']
            self.history_data['collapses'].union(c2o[source], c2o[target])

    def remove_group(self, index):                   # mesh.py:169-174
        if self.history_data is not None:
            hd = self.history_data
            old = hd['current2old'][index]
            hd['edges_mask'][-1][old] = 0
            hd['old2current'][old] = -1
            if self.export_folder:
                hd['collapses'].remove_group(old)

    def get_groups(self):                            # mesh.py:176-177
        self._pull_history()
        return self.history_data['groups'].pop()

    def get_occurrences(self):                       # mesh.py:179-180
        self._pull_history()
        return self.history_data['occurrences'].pop()

    def __clean_history(self, groups, pool_mask):    # mesh.py:182-192
        if self.history_data is not None:
            hd = self.history_data
            n = self._h['edges_count']
            alive = hd['old2current'] != -1
            hd['old2current'][alive] = np.arange(n, dtype=np.int32)
            hd['current2old'][0:n] = np.nonzero(alive)[0]
            if self.export_folder != '':
                hd['edges_mask'].append(hd['edges_mask'][-1].copy())
            hd['occurrences'].append(groups.get_occurrences())
            hd['groups'].append(groups.get_groups(pool_mask))
            hd['gemm_edges'].append(self._h['gemm_edges'].copy())
            hd['edges_count'].append(n)

    def unroll_gemm(self):                           # mesh.py:194-198
        self._own()
        hd = self.history_data
        hd['gemm_edges'].pop()
        self._h['gemm_edges'] = hd['gemm_edges'][-1]
        hd['edges_count'].pop()
        self._h['edges_count'] = hd['edges_count'][-1]

    def _pull_history(self):
        """Mirror what the device batch recorded for this mesh (one PoolRecord per MeshPool call) into the reference's
        history lists: dense clamped groups [ec_new, ec_old], un-clamped occurrences, pooled gemm_edges, edge counts, and the
        old2current / current2old maps.  Lazy: only host code that asks pays for the downloads."""
        bt = self._batch
        if bt is None or self.history_data is None:
            return
        hd, i = self.history_data, self._bi
        seen = getattr(self, '_hist_seen', 0)
        if seen >= len(bt.records):
            return
        bt.check()
        for rec in bt.records[seen:]:
            n_old = int(bt.levels[rec.level_in].ec[i].item())
            out = bt.levels[rec.level_out]
            n_new = int(out.ec[i].item())
            ptr = rec.rowptr[i, :n_new + 1].cpu().numpy()
            cols = rec.cols[i, :int(ptr[-1])].cpu().numpy()
            g = torch.zeros(n_new, n_old)
            g[np.repeat(np.arange(n_new), np.diff(ptr)), cols] = 1.0
            hd['groups'].append(g)
            hd['occurrences'].append(rec.occ[i, :n_old].cpu())
            hd['gemm_edges'].append(out.gemm[i, :n_new].cpu().numpy().astype(np.int64))
            hd['edges_count'].append(n_new)
            if rec.edge_mask is not None and 'old2current' in hd:
                keep = rec.edge_mask[i, :n_old].cpu().numpy().astype(bool)
                dead_old = hd['current2old'][:n_old][~keep]
                hd['old2current'][dead_old] = -1
                if not self.export_folder:                       # exporting meshes: MeshBatch.pool keeps the per-level masks
                    hd['edges_mask'][-1][dead_old] = False
                alive = hd['old2current'] != -1
                hd['old2current'][alive] = np.arange(n_new, dtype=np.int32)
                hd['current2old'][0:n_new] = np.nonzero(alive)[0]
        self._hist_seen = len(bt.records)

    # ---- OBJ export (mesh.py:74-124).  Host-side: the mirrored arrays are downloaded on demand. ----------
    def export(self, file=None, vcolor=None):
        """Write the current mesh as `v` / `f` / `e` records.  With no file name: `<export_folder>/<name>_<pool_count><ext>`,
        or nothing at all if the mesh has no export folder (mesh.py:74-80)."""
        if file is None:
            if not self.export_folder:
                return
            stem, ext = os.path.splitext(self.filename)
            file = '%s/%s_%d%s' % (self.export_folder, stem, self.pool_count, ext)
        v_mask = np.asarray(self.v_mask, dtype=bool)
        vs = np.asarray(self.vs)[v_mask]
        n = int(self.edges_count)
        edges = np.asarray(self.edges)[:n]
        sides = np.asarray(self.sides)[:n]
        left = np.array(np.asarray(self.gemm_edges)[:n], dtype=np.int64)       # consumed while the faces are walked
        renum = np.zeros(len(v_mask), dtype=np.int64)
        renum[v_mask] = np.arange(int(v_mask.sum()))
        faces = []
        for e0 in range(n):
            for start in (0, 2):
                if left[e0, start] == -1:
                    continue
                ring, key, side = [], e0, start
                for _ in range(3):                                               # walk the three edges of the face
                    nxt = int(left[key, side])
                    nside = int(sides[key, side])
                    nside += 1 - 2 * (nside % 2)
                    left[key, side] = -1
                    left[key, side + 1 - 2 * (side % 2)] = -1
                    key, side = nxt, nside
                    ring.append(key)
                face = []
                for i in range(3):
                    common = set(edges[ring[i]].tolist()) & set(edges[ring[(i + 1) % 3]].tolist())
                    face.append(int(renum[next(iter(common))]))
                faces.append(face)
        rows = []
        for vi, v in enumerate(vs):
            col = ' %f %f %f' % (vcolor[vi, 0], vcolor[vi, 1], vcolor[vi, 2]) if vcolor is not None else ''
            rows.append('v %f %f %f%s\n' % (v[0], v[1], v[2], col))
        for f in faces[:-1]:
            rows.append('f %d %d %d\n' % (f[0] + 1, f[1] + 1, f[2] + 1))
        rows.append('f %d %d %d' % (faces[-1][0] + 1, faces[-1][1] + 1, faces[-1][2] + 1))
        for a, b in edges:
            rows.append('\ne %d %d' % (renum[a] + 1, renum[b] + 1))
        with open(file, 'w+') as fh:
            fh.write(''.join(rows))

    def export_segments(self, segments):
        """Append the segment id of every edge to the `e` records of the files written by export(), one file per pool
        level; deeper levels keep the labels of the edges that survived (mesh.py:100-124)."""
        if not self.export_folder:
            return
        segments = np.asarray(segments)
        cur = segments
        masks = self.history_data.get('edges_mask', []) if self.history_data else []
        stem, ext = os.path.splitext(self.filename)
        for level in range(self.pool_count + 1):
            path = '%s/%s_%d%s' % (self.export_folder, stem, level, ext)
            out, k = [], 0
            with open(path) as fh:
                for line in fh:
                    if line[0] == 'e':
                        out.append('%s %d' % (line.strip(), cur[k]))
                        if k < len(cur):
                            k += 1
                            out.append('\n')
                    else:
                        out.append(line)
            with open(path, 'w') as fh:
                fh.write(''.join(out))
            if level < len(masks):
                m = np.asarray(masks[level], dtype=bool)
                cur = segments[:len(m)][m]


def _mirror(name):
    def get(self):
        self._sync()
        return self._h[name]

    def set_(self, v):
        self._sync()
        self._h[name] = v
    return property(get, set_)


for _n in Mesh._MIRRORED:
    setattr(Mesh, _n, _mirror(_n))
