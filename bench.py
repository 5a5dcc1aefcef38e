"""Benchmark of the MeshCNN per-edge hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cubes|shrec|humanseg] [--impl ours|reference]

One "step" = one full training step (zero_grad, forward incl. every MeshConv / MeshPool / MeshUnpool, loss, backward,
gradient allreduce when N > 1, Adam) over one batch of synthetic meshes.  Default workload = BASELINE.json configs[1]
(cubes: 64 x 750-edge meshes per GPU, ncf 64/128/256/256, pool_res 600/450/300/210, GroupNorm(16), resblocks 1).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]; scripts/cubes/train.sh of the reference
    'cubes': dict(arch='mconvnet', ncf=[64, 128, 256, 256], pool_res=[600, 450, 300, 210], E=750, B=64, nclasses=22,
                  resblocks=1, fc_n=100, groups=16, lr=2e-4, mesh=('ico', dict(nu=5)), mesh2=('torus', dict(m=10, n=25))),
    # configs[0]; scripts/shrec/train.sh
    'shrec': dict(arch='mconvnet', ncf=[64, 128, 256, 256], pool_res=[600, 450, 300, 180], E=750, B=16, nclasses=30,
                  resblocks=1, fc_n=100, groups=16, lr=2e-4, mesh=('ico', dict(nu=5)), mesh2=('torus', dict(m=10, n=25))),
    # configs[2]; scripts/human_seg/train.sh
    'humanseg': dict(arch='meshunet', ncf=[32, 64, 128, 256], pool_res=[1800, 1350, 600], E=2280, B=12, nclasses=8,
                     resblocks=3, lr=1e-3, mesh=('torus', dict(m=20, n=38)), mesh2=('torus', dict(m=20, n=38))),
    # configs[3]; scripts/coseg_seg/train.sh
    # configs[4]: MedMeshCNN-scale meshes (the reference cannot run them: its MeshUnion is a dense n x n matrix = 115 GB);
    # torus(238 x 238) = 169 932 edges padded to 170 000, one mesh per GPU, suggested deep pool chain
    'large': dict(arch='meshunet', ncf=[32, 64, 128, 256], pool_res=[128000, 96000, 48000], E=170000, B=1, nclasses=8,
                  resblocks=3, lr=1e-3, mesh=('torus', dict(m=238, n=238)), mesh2=('torus', dict(m=238, n=238)), dense=False),
    'coseg': dict(arch='meshunet', ncf=[32, 64, 128, 256], pool_res=[1800, 1350, 600], E=2280, B=32, nclasses=4,
                  resblocks=3, lr=1e-3, mesh=('torus', dict(m=20, n=38)), mesh2=('torus', dict(m=20, n=38))),
}


def make_inputs(cfg, B, first_seed):
    """Synthetic manifold meshes -> prepared data, standardised fp32 features [B, 5, E], labels."""
    from meshcnn_b200 import mesh_prepare, synth
    datas = []
    for i in range(B):
        s = first_seed + i
        kind, kw = cfg['mesh'] if s % 2 == 0 else cfg['mesh2']
        vs, faces = synth.make(kind, s, **kw)
        datas.append(mesh_prepare.from_arrays(vs, faces, filename='%s_%d.obj' % (kind, s)))
    E = cfg['E']
    x = np.zeros((B, 5, E), dtype=np.float64)
    for i, d in enumerate(datas):
        x[i, :, :d.edges_count] = d.features
    mean = x.mean(axis=(0, 2), keepdims=True)
    std = x.std(axis=(0, 2), keepdims=True)
    x = ((x - mean) / std).astype(np.float32)
    rs = np.random.RandomState(first_seed)
    if cfg['arch'] == 'mconvnet':
        labels = rs.randint(0, cfg['nclasses'], size=(B,)).astype(np.int64)
    else:
        labels = rs.randint(0, cfg['nclasses'], size=(B, E)).astype(np.int64)
        for i, d in enumerate(datas):
            labels[i, d.edges_count:] = -1
    return datas, x, labels


def build_net(cfg, impl):
    torch.manual_seed(0)
    if impl == 'ours':
        from meshcnn_b200 import networks as N
        if cfg['arch'] == 'mconvnet':
            net = N.MeshConvNet(N.get_norm_layer('group', cfg['groups']), 5, cfg['ncf'], cfg['nclasses'], cfg['E'],
                                cfg['pool_res'], cfg['fc_n'], cfg['resblocks'])
        else:
            net = N.MeshEncoderDecoder([cfg['E']] + cfg['pool_res'], [5] + cfg['ncf'], cfg['ncf'][::-1] + [cfg['nclasses']],
                                       blocks=cfg['resblocks'], transfer_data=True)
        N.init_weights(net, 'normal', 0.02)
    else:
        import functools
        from oracle import networks_oracle as ON
        from meshcnn_b200.networks import init_weights
        if cfg['arch'] == 'mconvnet':
            net = ON.OMeshConvNet(functools.partial(torch.nn.GroupNorm, cfg['groups']), 5, cfg['ncf'], cfg['nclasses'],
                                  cfg['E'], cfg['pool_res'], cfg['fc_n'], cfg['resblocks'], dense=cfg.get('dense', True))
        else:
            net = ON.OMeshEncoderDecoder([cfg['E']] + cfg['pool_res'], [5] + cfg['ncf'], cfg['ncf'][::-1] + [cfg['nclasses']],
                                         blocks=cfg['resblocks'], dense=cfg.get('dense', True))
        init_weights(net, 'normal', 0.02)
    return net


def loss_fn_for(cfg):
    return torch.nn.CrossEntropyLoss() if cfg['arch'] == 'mconvnet' else torch.nn.CrossEntropyLoss(ignore_index=-1)


# -----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (dense MeshUnion, per-mesh Python collapse, torch-CPU conv)
# -----------------------------------------------------------------------------------------------------------------
def cpu_port_meshes_per_s(cfg, sample_B, steps, warmup, budget_s=None):
    """`steps` timed steps after `warmup`; with `budget_s` the number of timed steps is chosen from the first timed step so
    that the sample is about that many seconds of CPU work (at least `steps`, at most 40)."""
    from oracle import meshcnn_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    datas, x, labels = make_inputs(cfg, sample_B, 0)
    net = build_net(cfg, 'reference').train()
    opt = torch.optim.Adam(net.parameters(), lr=cfg['lr'], betas=(0.9, 0.999))
    crit = loss_fn_for(cfg)
    xt, lt = torch.from_numpy(x), torch.from_numpy(labels)
    hold = cfg['arch'] == 'meshunet'
    times = []
    it = -1
    while True:
        it += 1
        if it >= warmup + steps:
            break
        meshes = [O.OracleMesh(d, hold_history=hold) for d in datas]       # fresh, outside the timed region (F11)
        xin = xt.clone().requires_grad_(True)
        t0 = time.perf_counter()
        opt.zero_grad()
        out = net(xin, meshes)
        loss = crit(out, lt)
        loss.backward()
        opt.step()
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
            if budget_s is not None and len(times) == 1:
                steps = max(steps, min(40, int(round(budget_s / max(times[0], 1e-3)))))
    return sample_B * len(times) / sum(times), threads, sum(times) / len(times)


# -----------------------------------------------------------------------------------------------------------------
# helpers
# -----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML from a background thread (every ~2 ms) while the timed
    region runs; falls back to one nvidia-smi query if NVML is unavailable."""
    REASONS = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'sw_power_cap': 0x4}

    def __init__(self, index):
        import threading
        self.sm, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                       # noqa: BLE001
            self.err = 'nvml: %s' % e
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def sample(self):
        """One sample, callable from the main thread between timed steps (the background thread can be starved of the GIL
        by a tight replay loop)."""
        if self._h is None:
            return False
        nv = self._nv
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for name, bit in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:                                       # noqa: BLE001
            self.err = 'nvml: %s' % e
            return False
        return True

    def _run(self):
        while not self._stop.is_set():
            if not self.sample():
                return
            time.sleep(0.01)

    def stop(self):
        self._stop.set()
        self._t.join(timeout=2)
        if not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'samples': 0, 'reasons': [self.err or 'no samples']}
        return {'sm_mhz': statistics.median(self.sm), 'sm_max_mhz': self.max_mhz, 'samples': len(self.sm),
                'reasons': sorted(self.reasons)}


def algorithmic_bytes(name, m):
    """SURVEY.md §8(d) per-launch algorithmic bytes (fp32 activations, int32 neighbour ids); see DESIGN.md."""
    if name == 'meshconv_fwd':
        return m['B'] * m['E'] * (4 * m['Cin'] + 4 * m['Cout'] + 16) + 4 * (5 * m['Cin'] * m['Cout'] + m['Cout'])
    if name == 'meshconv_bwd_data':
        return m['B'] * m['E'] * (4 * m['Cout'] + 8 * m['Cin'] + 16) + 20 * m['Cin'] * m['Cout']
    if name == 'meshconv_bwd_weight':
        return m['B'] * m['E'] * (4 * m['Cout'] + 4 * m['Cin'] + 16) + 20 * m['Cin'] * m['Cout']
    if name == 'pool_collapse':
        return 28 * m['B'] * (m['E_in'] + m['T'])
    if name == 'pool_norms':
        return 4 * m['B'] * m['E'] * (m['C'] + 1)
    if name in ('segmean_fwd', 'segmean_bwd'):                 # rows in + rows out + CSR (~1.07 nnz per input edge)
        return 4 * m['B'] * m['C'] * (m['E_in'] + m['T']) + 4 * m['B'] * (2 * m['E_in'] + m['T'])
    if name in ('unpool_fwd', 'unpool_bwd'):
        return 4 * m['B'] * m['C'] * (m['E_lo'] + m['E_hi']) + 4 * m['B'] * (3 * m['E_hi'])
    if name in ('norm_fwd', 'norm_bwd'):                       # x (+ y/dy) in, y/dx out
        n = m['nblocks'] * m['rows'] * m['C'] * 4
        return 2 * n if name == 'norm_fwd' else 4 * n
    if name == 'adam_flat':
        return 28 * m['n']
    if name in ('head_fwd', 'head_bwd'):                       # the activation read (fwd) / its gradient written (bwd) + both weight matrices
        return 4 * m['B'] * m['E'] * m['C'] + 4 * (m['C'] * m['H'] + m['H'] * m['K']) * (1 if name == 'head_fwd' else 2)
    return 0


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# -----------------------------------------------------------------------------------------------------------------
def reference_cpu(cfg, B, steps, warmup, budget_s=None):
    """The reference's own CPU implementation of the path, timed on this box's host cores: the UNMODIFIED reference
    (oracle/reference_arm.py: networks.define_classifier + Adam, fresh Mesh objects per step) on all B meshes of the workload
    -> kind "reference".  Only the 170 k-edge config, which the reference cannot run (dense n x n MeshUnion = 115 GB, SURVEY
    F15), falls back to the sparse oracle port -> kind "port"."""
    threads = os.cpu_count() or 1
    if cfg.get('dense', True):
        from oracle import reference_arm
        r = reference_arm.run(cfg, B, steps, warmup, budget_s=budget_s, threads=threads)
        split = {'fwd_s': r['fwd_s'], 'bwd_s': r['bwd_s'], 'opt_s': r['opt_s'], 'meshpool_s': r['pool_s'],
                 'meshpool_ms_per_mesh_per_level': 1e3 * r['pool_s'] / (B * len(cfg['pool_res'])),
                 'meshpool_share_of_step': r['pool_s'] / r['s_per_step']}
        sample = ('%d timed steps (+%d warm-up) of ALL %d meshes of the batch; unmodified reference (models/networks.py define_classifier, '
                  'torch.optim.Adam, fresh Mesh objects per step built outside the timed region), %d torch threads; the collapse '
                  'loop is single-threaded Python as in the reference; %.2f s/step' % (r['steps'], warmup, B, threads, r['s_per_step']))
        return r['meshes_per_s'], threads, r['s_per_step'], 'reference', sample, split, r['steps']
    v, threads, sec = cpu_port_meshes_per_s(cfg, B, steps, warmup, budget_s=budget_s)
    sample = ('timed steps of %d meshes; sparse oracle port (the reference cannot allocate its dense MeshUnion at this size); '
              '%.2f s/step' % (B, sec))
    return v, threads, sec, 'port', sample, None, steps


def run_reference(args, cfg, rank, world):
    """--impl reference: rank 0 only; same workload, metric and unit as our arm, every step over the whole batch."""
    if rank != 0:
        return None
    v, threads, sec, kind, sample, split, n = reference_cpu(cfg, cfg['B'], max(1, args.steps), max(0, args.warmup))
    line = {'impl': 'reference', 'metric': 'train meshes/sec (fwd+bwd+pool)', 'value': v, 'unit': 'meshes/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.config, cfg), 'global_batch': cfg['B'],
                       'sample': 'all %d meshes of the batch every step' % cfg['B']},
            'cpu_baseline': {'value': v, 'unit': 'meshes/s', 'cores': threads, 'kind': kind, 'sample': sample, 'split': split},
            'e2e': {'value': v, 'unit': 'meshes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    return line


def workload_name(name, cfg):
    if cfg['arch'] == 'mconvnet':
        return '%s: MeshConvNet ncf %s pool_res %s, %d-edge meshes, batch %d per GPU, GroupNorm(%d), resblocks %d' % (
            name, '/'.join(map(str, cfg['ncf'])), '/'.join(map(str, cfg['pool_res'])), cfg['E'], cfg['B'], cfg['groups'], cfg['resblocks'])
    return '%s: MeshEncoderDecoder ncf %s pool_res %s, %d-edge meshes, batch %d per GPU, resblocks %d' % (
        name, '/'.join(map(str, cfg['ncf'])), '/'.join(map(str, cfg['pool_res'])), cfg['E'], cfg['B'], cfg['resblocks'])


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices) also write to fd 1, so
    the real stdout is set aside for the result and fd 1 is pointed at stderr for everything else."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, 'w')


def _finish(out, line, cleanup=None):
    """Print the result, release the CUDA graphs and the NCCL communicator in that order, and leave through the normal
    interpreter exit (atexit hooks run, so the driver can record which .so files this process mapped).  Graphs that captured
    a collective must be destroyed BEFORE their communicator; a watchdog ends the process if teardown still blocks."""
    if line is not None:
        out.write(json.dumps(line) + '\n')
        out.flush()
    sys.stderr.flush()
    import threading

    def _bail():
        sys.stderr.write('bench.py: teardown did not finish within 90 s; forcing exit\n')
        sys.stderr.flush()
        os._exit(0)
    wd = threading.Timer(90.0, _bail)
    wd.daemon = True
    wd.start()
    if cleanup is not None:
        cleanup()
    sys.exit(0)


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--config', default='cubes', choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch size')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-flush', action='store_true', help='diagnostics: no L2 flush between timed steps (the number is then not a bench value)')
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying a CUDA graph')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'tf32'],
                    help="arithmetic of the tensor-core MeshConv kernels: fp32 = 3xTF32 (the headline), tf32 = one TF32 MMA per product "
                         "(opt-in reduced precision, inside the 2e-2 tolerance BASELINE.json allows; reported beside, never instead of, fp32)")
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: the config batch PER GPU; strong: the config batch is the GLOBAL batch, sharded over the ranks '
                         '(uneven shards are weighted by mesh count in the gradient all-reduce)')
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg['B'] = args.batch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        _finish(out, run_reference(args, cfg, rank, world))
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from meshcnn_b200 import _lib, ddp, profiler
    from meshcnn_b200 import functional as Fn
    Fn.set_precision(args.precision)
    from meshcnn_b200.mesh import Mesh, MeshBatch

    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    B, E = cfg['B'], cfg['E']
    B_global = world * B
    if args.scaling == 'strong':                                        # rank r owns a contiguous slice of the batch
        B_global = cfg['B']
        if B_global < world:
            raise SystemExit('strong scaling needs at least one mesh per rank (batch %d, %d ranks)' % (B_global, world))
        lo, hi = ddp.shard_range(B_global, rank, world)
        datas, x, labels = make_inputs(cfg, B_global, 0)                # every rank builds the global batch, keeps its slice
        datas, x, labels = datas[lo:hi], np.ascontiguousarray(x[lo:hi]), np.ascontiguousarray(labels[lo:hi])
        B = hi - lo
    else:
        datas, x, labels = make_inputs(cfg, B, rank * B)
    net = build_net(cfg, 'ours').to(dev).train()
    if world > 1:
        ddp.broadcast_parameters(net)
    # grads are views into one flat buffer; uneven shards: each rank's mean gradient is weighted by its share of the batch
    reducer = ddp.FlatGradReducer(net, local_items=B, global_items=B_global) if args.scaling == 'strong' else ddp.FlatGradReducer(net)
    from meshcnn_b200.optim import FlatAdam
    opt = FlatAdam(reducer, lr=cfg['lr'], betas=(0.9, 0.999))            # one kernel over the flat parameter buffer
    crit = loss_fn_for(cfg)
    hold = cfg['arch'] == 'meshunet'

    # ---- device-resident arm ---------------------------------------------------------------------------------------
    from meshcnn_b200.graphed import GraphedTrainStep
    meshes = [Mesh(data=d, hold_history=hold) for d in datas]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    mode = 'eager' if args.no_graph else 'cuda-graph'
    gs = None
    launches_per_step = None
    if not args.no_graph:
        try:
            l0 = _lib.launch_count()
            gs = GraphedTrainStep(net, crit, reducer, opt, meshes, x, labels, E, source='device')
            launches_per_step = (_lib.launch_count() - l0) // (gs.warmup_steps + 1)
        except Exception as e:                                       # noqa: BLE001  (e.g. a collective that cannot be captured)
            sys.stderr.write('CUDA-graph capture failed (%s); running eagerly\n' % (e,))
            gs, mode = None, 'eager (graph capture failed)'
            torch.cuda.synchronize()
    if gs is not None:
        batch = gs.batch
        step = gs.run
    else:
        batch = MeshBatch(meshes, E, dev)
        batch.snapshot()
        x_dev = torch.from_numpy(x).to(dev).requires_grad_(True)       # the reference trains with requires_grad inputs
        y_dev = torch.from_numpy(labels).to(dev)

        def step():
            batch.reset()
            reducer.zero()
            x_dev.grad = None
            out = net(x_dev, meshes)
            loss = crit(out, y_dev)
            loss.backward()
            reducer.allreduce()
            opt.step()
            return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    batch.check()
    # the clock sampler (NVML init: a few ms on rank 0 only) must exist BEFORE the barrier: started after it, rank 0 entered the
    # timed region late and every other rank's first timed step waited ~4 ms for it at the all-reduce -- 0.2-0.3 ms per step
    # of a 20-step run, which is what rounds 1 and 2 had been reading as a multi-GPU scaling loss
    sampler = ClockSampler(local) if rank == 0 else None
    import gc
    gc.collect()                                                        # (before the barrier, like everything that takes time)
    gc.disable()                                                        # a collector pause on one rank stalls every rank at the all-reduce
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    evs = []
    for _ in range(args.steps):
        if not args.no_flush:
            flush.fill_(1)                                              # L2 flush between timed steps (not timed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
        if sampler is not None and (len(evs) & 3) == 0:
            sampler.sample()                                            # GPU is busy with this step; the call is not timed
    torch.cuda.synchronize()
    gc.enable()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler is not None else None
    launches = (_lib.launch_count() - launches0) if gs is None else launches_per_step * args.steps
    batch.check()
    per_step = sorted(a.elapsed_time(b) for a, b in evs)
    total_ms = sum(per_step)
    spread = torch.tensor([per_step[0], per_step[len(per_step) // 2], per_step[-1], total_ms / len(per_step)], dtype=torch.float64, device=dev)
    spreads = [spread]
    if world > 1:
        spreads = [torch.zeros_like(spread) for _ in range(world)]
        dist.all_gather(spreads, spread)
    step_spread = [{'min': float(t[0]), 'median': float(t[1]), 'max': float(t[2]), 'mean': float(t[3])} for t in spreads]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B_global * args.steps / (total_ms / 1e3)

    # ---- per-kernel CUDA-event timings: the same step launched kernel by kernel (events cannot bracket the nodes of a
    # replayed graph), L2 flushed between steps as above -------------------------------------------------------------
    pmeshes = [Mesh(data=d, hold_history=hold) for d in datas]
    pbatch = MeshBatch(pmeshes, E, dev)
    pbatch.snapshot()
    px = torch.from_numpy(x).to(dev).requires_grad_(True)
    py = torch.from_numpy(labels).to(dev)

    def eager_step():
        pbatch.reset()
        reducer.zero()
        px.grad = None
        loss = crit(net(px, pmeshes), py)
        loss.backward()
        reducer.allreduce()
        opt.step()

    for _ in range(2):
        eager_step()
    torch.cuda.synchronize()
    k_prof = min(args.steps, 10)
    profiler.enabled = True
    profiler.reset()
    pe = []
    for _ in range(k_prof):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eager_step()
        b.record()
        pe.append((a, b))
    profiler.enabled = False
    recs = profiler.collect()
    eager_ms = sum(a.elapsed_time(b) for a, b in pe)
    del pbatch, pmeshes
    # where the multi-GPU step loses time, measured on the device: the same step captured once more WITHOUT the all-reduce
    # node and replayed on every rank by itself (no collective: ranks do not wait for each other) gives each rank's own
    # critical path; the timed region above gives the step with the collective.
    attribution = None
    if world > 1 and gs is not None:
        reducer.enabled = False
        saved_state = gs._training_state()                            # these replays update with un-reduced gradients: undone below
        own, err = float('nan'), None
        try:
            gs0 = GraphedTrainStep(net, crit, reducer, opt, [Mesh(data=d, hold_history=hold) for d in datas], x, labels, E, source='device')
            for _ in range(3):
                gs0.run()
            torch.cuda.synchronize()
            ev0 = []
            for _ in range(args.steps):
                flush.fill_(1)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                gs0.run()
                eb.record()
                ev0.append((ea, eb))
            torch.cuda.synchronize()
            own = sum(ea.elapsed_time(eb) for ea, eb in ev0) / len(ev0)
            gs0.close()
            del gs0
        except Exception as e:                                       # noqa: BLE001  (diagnostics must not cost the bench line)
            err = str(e)[:200]
        finally:
            reducer.enabled = True
            gs._restore_training_state(saved_state)
            torch.cuda.synchronize()
        mine = torch.tensor([own], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        own_r = [float(t[0]) for t in allr]
        step_ms = total_ms / args.steps
        attribution = {'error': err} if any(v != v for v in own_r) else {'own_step_ms_per_rank': own_r, 'step_ms_with_allreduce': step_ms,
                       'slowest_rank_ms': max(own_r), 'mean_rank_ms': sum(own_r) / world,
                       'straggler_ms': max(own_r) - sum(own_r) / world,
                       'allreduce_and_skew_ms': step_ms - max(own_r),
                       'allreduce_and_skew_ms_by_median_step': max(r['median'] for r in step_spread) - max(own_r),
                       'note': 'device time (CUDA events) of the step replayed as a graph: per rank without the all-reduce node '
                               '(ranks run alone) vs the timed region with it.  straggler = slowest rank - mean rank (every '
                               'rank has other meshes: its pool kernels last as long as its slowest mesh); the rest of the '
                               'difference to the timed step is the NCCL kernel and the arrival skew it absorbs'}

    # ---- end-to-end arm: host buffers in, loss out, every step ------------------------------------------------------
    e2e = None
    ge_holder = [None]
    if not args.no_e2e:
        k_e2e = args.steps
        n_fresh = min(k_e2e + 2, 12)
        fresh = [[Mesh(data=d, hold_history=hold) for d in datas] for _ in range(n_fresh)]   # the data loader's job (F11/F12)
        t_pack = [0.0]
        if args.no_graph or gs is None:
            x_pin = torch.from_numpy(x).pin_memory()
            y_pin = torch.from_numpy(labels).pin_memory()
            h2d = [0]

            def e2e_step(i):
                ms = [Mesh(data=d, hold_history=hold) for d in datas] if i >= n_fresh else fresh[i]
                tp = time.perf_counter()
                bt = MeshBatch(ms, E, dev)                               # packs + uploads the topology blob (pinned)
                t_pack[0] += time.perf_counter() - tp
                xin = x_pin.to(dev, non_blocking=True).requires_grad_(True)
                yin = y_pin.to(dev, non_blocking=True)
                h2d[0] = bt.blob.numel() + x_pin.numel() * 4 + y_pin.numel() * 8
                reducer.zero()
                loss = crit(net(xin, ms), yin)
                loss.backward()
                reducer.allreduce()
                opt.step()
                return float(loss.item())                                # D2H read of the step's result
            h2d_bytes = lambda: h2d[0]
        else:
            from meshcnn_b200.graphed import PipelinedTrainStep
            ge = PipelinedTrainStep(net, crit, reducer, opt, fresh[0], x, labels, E)
            ge_holder[0] = ge
            x_np, y_np = torch.from_numpy(x), torch.from_numpy(labels)
            pending = [None]

            def e2e_step(i):
                """Step i is packed on the host and submitted while the GPU still runs step i-1; then the loss of step i-1 is
                read (D2H copy inside its graph + one event wait).  Every step's loss is read inside the timed region."""
                ms = fresh[i % n_fresh]
                for m in ms:                                             # a used Mesh is single-use (F11): hand the loader's
                    m._batch = None                                      # pristine host copy to the step again
                tp = time.perf_counter()
                slot = ge.submit(ms, x_np, y_np)                         # pack into pinned buffers, replay (H2D copies are graph nodes)
                t_pack[0] += time.perf_counter() - tp
                prev, pending[0] = pending[0], slot
                return None if prev is None else prev.result()
            h2d_bytes = lambda: ge.h2d_bytes

        for i in range(2):
            e2e_step(i)
        torch.cuda.synchronize()
        t_pack[0] = 0.0
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(k_e2e):
            e2e_step(2 + i)
        if not (args.no_graph or gs is None):
            pending[0].result()                                          # the last step's loss
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {'value': B_global * k_e2e / float(t.item()), 'unit': 'meshes/s', 'h2d_bytes_per_step': int(h2d_bytes()),
               'd2h_bytes_per_step': 4, 'timing': 'host wall clock between device synchronisations, max over ranks',
               'host_pack_ms_per_step': 1e3 * t_pack[0] / k_e2e,
               'api': 'meshcnn_b200.graphed.PipelinedTrainStep.submit(meshes, x, y) from host buffers, loss read back every step' if gs is not None else
                      'net(x, meshes) eager from host buffers'}

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    def cleanup():
        """CUDA graphs first (they hold the captured NCCL kernels), then the communicator."""
        import gc
        for g in [gs] + (list(ge_holder[0].slots) if ge_holder[0] is not None else []):
            if g is not None:
                g.close()
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.destroy_process_group()

    if rank != 0:
        _finish(out, None, cleanup)

    # ---- roofline: the dominant kernel by time, whichever it is, + one entry per kernel family ------------------------
    agg = {}
    for name, meta, ms in recs:
        key = (name, tuple(sorted(meta.items())))
        tot, cnt = agg.get(key, (0.0, 0))
        agg[key] = (tot + ms, cnt + 1)
    by_kernel = {}
    for (name, meta), (tot, cnt) in agg.items():
        by_kernel[name] = by_kernel.get(name, 0.0) + tot
    hbm_peak, hbm_src = peaks()
    mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    tf32_peak = json.load(open(mp))['bf16_tflops_sustained'] / 2.0 if os.path.exists(mp) else 1125.0
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    traffic_db = json.load(open(tpath)) if os.path.exists(tpath) else {}
    timing_note = ('EAGER attribution pass: CUDA events around each C-ABI call in %d steps launched kernel by kernel (%.3f ms/step) '
                   'right after the timed region -- events cannot bracket nodes of the replayed graph (%.3f ms/step); shares of '
                   'the graph step are in profiles/ (ncu launch list)' % (k_prof, eager_ms / k_prof, total_ms / args.steps))

    def family_entry(fam):
        """Roofline of the heaviest launch shape of one kernel family."""
        (name, meta), (tot, cnt) = max(((k, v) for k, v in agg.items() if k[0] == fam), key=lambda kv: kv[1][0])
        m = dict(meta)
        avg_ms = tot / cnt
        base = name[:-3] if name.endswith('_tc') else name
        nbytes = algorithmic_bytes(base, m)
        ent = traffic_db.get('%s|%s' % (name, ','.join('%s=%s' % kv for kv in sorted(m.items()))))
        e = {'kernel': name, 'shape': m, 'avg_launch_ms': avg_ms, 'launches_per_step': sum(c for k, (t, c) in agg.items() if k[0] == fam) / k_prof,
             'share_of_eager_step': by_kernel[fam] / eager_ms, 'algorithmic_bytes': nbytes,
             'traffic': ent['dram_bytes'] if ent else None,
             'traffic_source': ('%s row %d (ncu --set full, one launch)' % (ent.get('file', 'profiles/r01_v11_ncu_top_kernels.md'), ent['row'])) if ent else None}
        if name.endswith('_tc'):
            # tensor-core MeshConv: algorithmic flops of the contraction (fp32 result); 3 TF32 MMAs are issued per algorithmic
            # MMA (3xTF32 split), so the tensor pipe is busy for 3x the achieved figure.  Peak = TF32 dense = bf16 sustained / 2.
            flops = 2.0 * m['B'] * m['E'] * 5 * m['Cin'] * m['Cout']
            ach = flops / (avg_ms * 1e-3) / 1e12
            e.update(bound='tensor', achieved=ach, peak=tf32_peak, unit='TFLOP/s', frac=ach / tf32_peak, algorithmic_flops=flops,
                     mma_passes=3, tensor_pipe_frac=3.0 * ach / tf32_peak, hbm_GBps=nbytes / (avg_ms * 1e-3) / 1e9,
                     peak_source='TF32 dense = measured bf16 sustained (MEASURED_PEAKS.json) / 2')
        else:
            ach = nbytes / (avg_ms * 1e-3) / 1e9
            e.update(bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s', frac=ach / hbm_peak, peak_source=hbm_src)
        if name == 'pool_collapse':
            ncol = max(1, (m['E_in'] - m['T']) // 3)
            e.update(limiter='latency: a dependent chain of edge collapses per mesh (one CTA per mesh); the HBM fraction is reported '
                             'for completeness, us/collapse is the figure of merit (SURVEY 8d)',
                     collapses_per_mesh=ncol, us_per_collapse=1e3 * avg_ms / ncol, us_per_mesh_per_level=1e3 * avg_ms / m['B'])
        return e

    roofline, pool_info = None, None
    if agg:
        fams = sorted(by_kernel, key=by_kernel.get, reverse=True)
        entries = {f: e for f, e in ((f, family_entry(f)) for f in fams) if e['algorithmic_bytes'] > 0}
        dominant = fams[0]
        top = entries.get(dominant) or next(iter(entries.values()))
        roofline = dict(top, dominant_by_time=dominant, timing=timing_note,
                        share_of_step={k: v / eager_ms for k, v in sorted(by_kernel.items())},
                        kernels=[entries[f] for f in fams if f in entries and entries[f] is not top])
        pool_info = entries.get('pool_collapse')

    cpu = None
    if not args.no_cpu_baseline and world == 1:                  # rank 0 at N = 1 only; a bounded sample (~15 s of CPU work)
        v, threads, sec, kind, sample, split, n = reference_cpu(cfg, B, 2, 1, budget_s=12.0)
        cpu = {'value': v, 'unit': 'meshes/s', 'cores': threads, 'kind': kind, 'sample': sample, 'split': split}
        if split is not None and pool_info is not None:
            # MeshPool alone (SURVEY 8d asks for >= 50x at one GPU): the reference's MeshPool.forward wall time per step
            # against the sum of our pool kernels (norms + collapse + segment mean) per step
            ours_ms = sum(by_kernel.get(k, 0.0) for k in ('pool_collapse', 'pool_norms', 'segmean_fwd')) / k_prof
            cpu['meshpool_speedup'] = {'reference_ms_per_step': 1e3 * split['meshpool_s'], 'ours_ms_per_step': ours_ms,
                                       'ratio': 1e3 * split['meshpool_s'] / max(ours_ms, 1e-9),
                                       'note': 'reference: wall time inside MeshPool.forward (host); ours: CUDA-event time of the pool kernels in the eager attribution pass'}

    line = {'metric': 'train meshes/sec (fwd+bwd+pool)', 'value': value, 'unit': 'meshes/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'ms_per_step_by_rank': step_spread,
            'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'tf32 (single-pass TF32 MMAs in the MeshConv kernels, fp32 everywhere else; opt-in)', 'data': 'synthetic',
            'config': {'workload': workload_name(args.config, cfg) + (' [GLOBAL batch, sharded]' if args.scaling == 'strong' else ''),
                       'global_batch': B_global,
                       'parallelism': 'dp%d (meshes sharded, one flat NCCL gradient allreduce)' % world,
                       'l2': ('flushed between timed steps (256 MiB write, untimed)' if not args.no_flush else 'NOT flushed (diagnostic run)') + '; per-step CUDA events, max over ranks',
                       'optimizer': 'Adam (in the timed region)', 'launch': mode},
            'roofline': roofline, 'pool_collapse': pool_info, 'cpu_baseline': cpu, 'e2e': e2e, 'multi_gpu_attribution': attribution, 'gpu_launches': int(launches), 'clocks': clocks}
    _finish(out, line, cleanup)


if __name__ == '__main__':
    main()
