"""autograd.Function wrappers over the C ABI.  All tensors here are edge-major [B, E, C] fp32 on a CUDA device;
every call enqueues on the current stream and never synchronises."""
import os

import torch

from . import _lib
from . import profiler

# tensor-core (tcgen05, 3xTF32) path for the layers that are real dense GEMMs; MESHCNN_B200_TC=0 forces the FFMA path
USE_TC = os.environ.get('MESHCNN_B200_TC', '1') != '0'
# run the weight-gradient kernel of a MeshConv on a side stream, concurrently with its data-gradient kernel (both only
# read dout / x): fills the tail waves of either kernel.  MESHCNN_B200_BWD_OVERLAP=0 keeps them on one stream.
BWD_OVERLAP = os.environ.get('MESHCNN_B200_BWD_OVERLAP', '1') != '0'
# persistent ("stream-K") forward kernel: (tile, K block) units spread evenly over the SMs; MESHCNN_B200_TC_SK=0 launches
# one CTA per tile instead
USE_SK = os.environ.get('MESHCNN_B200_TC_SK', '1') != '0'
# inside a step opened with begin_step() (gradients zeroed by the caller, every parameter used once per step) the weight /
# bias / gamma / beta gradient kernels write straight into p.grad (the flat gradient bucket) and autograd gets None for them:
# no AccumulateGrad add kernel per parameter (32 tiny launches per step of the cubes net).  MESHCNN_B200_DIRECT_GRADS=0 disables.
DIRECT_GRADS = os.environ.get('MESHCNN_B200_DIRECT_GRADS', '1') != '0'
# Arithmetic of the tensor-core MeshConv kernels.  'fp32' (default): every product as three TF32 MMAs, fp32 accuracy (1e-6).
# 'tf32': ONE TF32 MMA per product -- operands rounded to 10 mantissa bits, fp32 accumulation, ~5e-4 relative error, well inside
# the 2e-2 that BASELINE.json allows a reduced-precision path; a third of the tensor work.  Opt-in: MESHCNN_B200_PRECISION=tf32
# or set_precision('tf32'); never the default, and bench.py reports it beside the fp32 number, not instead of it.
PRECISION = _lib.PREC_TF32 if os.environ.get('MESHCNN_B200_PRECISION', 'fp32').lower() == 'tf32' else _lib.PREC_FP32


def set_precision(name):
    """'fp32' (3xTF32, the default) or 'tf32' (single-pass TF32) for the tensor-core MeshConv kernels of layers called from now on."""
    global PRECISION
    if name not in ('fp32', 'tf32'):
        raise ValueError("precision must be 'fp32' or 'tf32', got %r" % (name,))
    PRECISION = _lib.PREC_TF32 if name == 'tf32' else _lib.PREC_FP32


_SIDE = {}
_SK_WS = {}
STEP_TOKEN = -1          # >= 0 while a step prepared with begin_step() is being enqueued
GRAD_REDUCER = None      # ddp.FlatGradReducer that wants to start its tail all-reduce from inside backward (multi-GPU steps)
_NEXT_TOKEN = 0


def begin_step(net):
    """Start of a training step: build the tensor-core weight images of every MeshConv on the side stream, off the critical
    path (they depend on the weights only).  Returns the step token; call end_step() after backward."""
    global STEP_TOKEN, _NEXT_TOKEN
    from .mesh_conv import MeshConv
    _NEXT_TOKEN += 1
    STEP_TOKEN = _NEXT_TOKEN
    dev = next(net.parameters()).device
    main, side = torch.cuda.current_stream(), _side_stream(dev)
    side.wait_stream(main)
    for m in net.modules():
        if isinstance(m, MeshConv):
            m.prepare_images(STEP_TOKEN, side)
    return STEP_TOKEN


def end_step():
    """End of the step started by begin_step(): join the side stream, images are no longer valid."""
    global STEP_TOKEN, _NEXT_TOKEN
    for s in _SIDE.values():
        torch.cuda.current_stream().wait_stream(s)
    STEP_TOKEN = -1


def _side_stream(device):
    s = _SIDE.get(device)
    if s is None:
        s = _SIDE[device] = torch.cuda.Stream(device=device)
    return s


def _sk_workspace(device):
    """Partial-tile slots + flags of the stream-K forward kernel, one per (device, stream): launches on one stream are
    ordered, so they can share it.  Zeroed once here; the kernel leaves every flag lowered."""
    # inside a step opened with begin_step() (the graphed train step: warm-up, capture and replays are ordered among
    # themselves) one workspace per device, allocated during the warm-up -- an allocation during capture would put its 50 MB
    # zero-fill into the graph and replay it every step
    key = (device, 'step') if STEP_TOKEN >= 0 else (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _SK_WS.get(key)
    if ws is None:
        ws = _SK_WS[key] = torch.zeros(int(_lib.lib().mcnn_meshconv_fwd_tc_sk_ws()), dtype=torch.uint8, device=device)
    return ws


def _direct_grad(*params):
    """The .grad buffers of `params` if the kernels may write into them directly (see DIRECT_GRADS), else None."""
    if not (DIRECT_GRADS and STEP_TOKEN >= 0):
        return None
    out = []
    for p in params:
        if p is None:
            out.append(None)
            continue
        g = getattr(p, 'grad', None)
        if not isinstance(p, torch.nn.Parameter) or g is None or not g.is_contiguous() or g.dtype != torch.float32 or g.shape != p.shape:
            return None
        out.append(g)
    return out


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('meshcnn_b200 layers run on CUDA only (no CPU fallback); got a %s tensor' % t.device)


class MeshConvFn(torch.autograd.Function):
    """Fused MeshConv (mesh_conv.py:20-26) with the backward of SURVEY.md Appendix A."""

    @staticmethod
    def forward(ctx, x, weight, bias, level, wt_ws, img=None, pass_input=False):
        """pass_input=True returns (out, x): the caller uses the second output wherever it needs x again (the residual branch
        of DownConv / UpConv, networks.py:222-233 / :276-288).  Both gradients of x then arrive HERE, and the data-gradient
        kernel scatters straight into the residual's gradient instead of into zeros: no zero fill and no gradient add."""
        _chk_cuda(x, weight, bias)
        x_in = x
        B, E, Cin = x.shape
        Cout = weight.shape[0]
        x = x.contiguous()
        w = weight.contiguous()
        out = torch.empty(B, E, Cout, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        use_tc = USE_TC and L.mcnn_meshconv_tc_supported(Cin, Cout)
        # the tensor-core forward also keeps sign(f1 - f3), sign(f2 - f4) for the data-gradient kernel when x needs a gradient
        signs = None
        if use_tc and ctx.needs_input_grad[0] and Cout % 32 == 0:
            signs = torch.empty(2, B * E, Cin, dtype=torch.int8, device=x.device)
        with profiler.span('meshconv_fwd_tc' if use_tc else 'meshconv_fwd', B=B, E=E, Cin=Cin, Cout=Cout):
            if use_tc:
                if img is not None:                          # weight image prepared ahead of time on the side stream
                    torch.cuda.current_stream().wait_event(img[2])
                    wptr, iptr = None, img[0].data_ptr()
                else:
                    wptr, iptr = w.data_ptr(), wt_ws.data_ptr()
                if USE_SK:
                    _lib.check(L.mcnn_meshconv_fwd_tc_sk_p(x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(), wptr,
                                                           _lib.ptr(bias), out.data_ptr(), iptr, _lib.ptr(signs),
                                                           _sk_workspace(x.device).data_ptr(), B, E, Cin, Cout, PRECISION,
                                                           _lib.current_stream()), 'mcnn_meshconv_fwd_tc_sk')
                else:
                    _lib.check(L.mcnn_meshconv_fwd_tc_signs(x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(), wptr,
                                                            _lib.ptr(bias), out.data_ptr(), iptr, _lib.ptr(signs), B, E, Cin, Cout,
                                                            _lib.current_stream()), 'mcnn_meshconv_fwd_tc')
            else:
                _lib.check(L.mcnn_meshconv_fwd(x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(), w.data_ptr(),
                                               _lib.ptr(bias), out.data_ptr(), wt_ws.data_ptr(), B, E, Cin, Cout,
                                               _lib.current_stream()), 'mcnn_meshconv_fwd')
        ctx.save_for_backward(x, w)
        ctx.prec = PRECISION
        ctx.params = (weight, bias)
        ctx.signs = signs
        ctx.img = img if use_tc else None
        ctx.level = level
        ctx.has_bias = bias is not None
        ctx.pass_input = bool(pass_input)
        if pass_input:
            return out, x_in
        return out

    @staticmethod
    def backward(ctx, dout, dpass=None):
        x, w = ctx.saved_tensors
        level = ctx.level
        B, E, Cin = x.shape
        Cout = w.shape[0]
        dout = dout.contiguous()
        L = _lib.lib()
        st = _lib.current_stream()
        dx = dw = db = None
        tc_ok = USE_TC and Cin % 32 == 0
        want_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        fork = BWD_OVERLAP and not profiler.enabled and ctx.needs_input_grad[0] and want_w
        direct = _direct_grad(*ctx.params) if (want_w and ctx.needs_input_grad[1] and (not ctx.has_bias or ctx.needs_input_grad[2])) else None
        if fork:
            # every buffer is allocated on the main stream; the side stream only runs the dW kernels between two events
            use_tc_w = tc_ok and Cout % 4 == 0
            if direct is not None:                           # zeroed by the step; the kernels overwrite (tc) or add into zero (FFMA)
                dw, db = direct
            else:
                dw = torch.empty_like(w) if use_tc_w else torch.zeros_like(w)
                db = None
                if ctx.has_bias:
                    db = torch.empty(Cout, dtype=torch.float32, device=x.device) if use_tc_w else torch.zeros(Cout, dtype=torch.float32, device=x.device)
            ws = torch.empty(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, Cin, Cout), dtype=torch.float32, device=x.device) if use_tc_w else None
            main, side = torch.cuda.current_stream(), _side_stream(x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if use_tc_w:
                    _lib.check(L.mcnn_meshconv_bwd_weight_tc_p(dout.data_ptr(), x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(),
                                                               dw.data_ptr(), _lib.ptr(db), ws.data_ptr(), B, E, Cin, Cout, ctx.prec,
                                                               side.cuda_stream), 'mcnn_meshconv_bwd_weight_tc')
                else:
                    _lib.check(L.mcnn_meshconv_bwd_weight(dout.data_ptr(), x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(),
                                                          dw.data_ptr(), _lib.ptr(db), B, E, Cin, Cout, side.cuda_stream),
                               'mcnn_meshconv_bwd_weight')
        if ctx.needs_input_grad[0]:
            # the gradient that reached x through the pass-through output: scatter into it (it is this node's to keep: the norm
            # block allocated it for this edge alone); anything unusual is added afterwards instead
            seed = dpass if (dpass is not None and dpass.dtype == torch.float32 and dpass.shape == x.shape and dpass.is_contiguous()
                             and dpass._base is None and dpass.data_ptr() != dout.data_ptr()) else None
            dx = seed if seed is not None else torch.zeros_like(x)
            if tc_ok and Cout % 32 == 0:
                pre = ctx.img is not None
                wtb = ctx.img[1] if pre else torch.empty(2 * 5 * Cin, Cout, dtype=torch.float32, device=x.device)
                with profiler.span('meshconv_bwd_data_tc', B=B, E=E, Cin=Cin, Cout=Cout):
                    _lib.check(L.mcnn_meshconv_bwd_data_tc_signs_p(dout.data_ptr(), x.data_ptr(), _lib.ptr(ctx.signs), level.gemm.data_ptr(),
                                                                   level.ec.data_ptr(), None if pre else w.data_ptr(), dx.data_ptr(),
                                                                   wtb.data_ptr(), B, E, Cin, Cout, ctx.prec, st), 'mcnn_meshconv_bwd_data_tc')
            else:
                with profiler.span('meshconv_bwd_data', B=B, E=E, Cin=Cin, Cout=Cout):
                    _lib.check(L.mcnn_meshconv_bwd_data(dout.data_ptr(), x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(),
                                                        w.data_ptr(), dx.data_ptr(), B, E, Cin, Cout, st), 'mcnn_meshconv_bwd_data')
        if fork:
            main.wait_stream(side)
        elif want_w:
            if tc_ok and Cout % 4 == 0:
                if direct is not None:
                    dw, db = direct
                else:
                    dw = torch.empty_like(w)
                    db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
                ws = torch.empty(L.mcnn_meshconv_bwd_weight_tc_ws(B, E, Cin, Cout), dtype=torch.float32, device=x.device)
                with profiler.span('meshconv_bwd_weight_tc', B=B, E=E, Cin=Cin, Cout=Cout):
                    _lib.check(L.mcnn_meshconv_bwd_weight_tc_p(dout.data_ptr(), x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(),
                                                               dw.data_ptr(), _lib.ptr(db), ws.data_ptr(), B, E, Cin, Cout, ctx.prec, st),
                               'mcnn_meshconv_bwd_weight_tc')
            else:
                if direct is not None:
                    dw, db = direct
                else:
                    dw = torch.zeros_like(w)
                    db = torch.zeros(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
                with profiler.span('meshconv_bwd_weight', B=B, E=E, Cin=Cin, Cout=Cout):
                    _lib.check(L.mcnn_meshconv_bwd_weight(dout.data_ptr(), x.data_ptr(), level.gemm.data_ptr(), level.ec.data_ptr(),
                                                          dw.data_ptr(), _lib.ptr(db), B, E, Cin, Cout, st), 'mcnn_meshconv_bwd_weight')
        if direct is not None:
            dw = db = None                                   # already in weight.grad / bias.grad
            if GRAD_REDUCER is not None and want_w:
                GRAD_REDUCER.conv_backward_launched(ctx.params[0])
        if dpass is not None and dx is not None and dx is not dpass:
            dx = dx + dpass
        elif dpass is not None and dx is None and ctx.needs_input_grad[0]:
            dx = dpass
        return dx, dw, db, None, None, None, None


def pool_norms(x):
    """fp32 squared feature norms [B, E] (mesh_pool.py:186); not differentiable by construction."""
    B, E, C = x.shape
    norms = torch.empty(B, E, dtype=torch.float32, device=x.device)
    with profiler.span('pool_norms', B=B, E=E, C=C):
        _lib.check(_lib.lib().mcnn_pool_norms(x.data_ptr(), norms.data_ptr(), B, E, C, _lib.current_stream()), 'mcnn_pool_norms')
    return norms


class SegMeanFn(torch.autograd.Function):
    """rebuild_features_average (mesh_union.py:27-36) over the CSR groups of a PoolRecord."""

    @staticmethod
    def forward(ctx, x, rec, ec_in, ec_out):
        _chk_cuda(x)
        x = x.contiguous()
        B, E_in, C = x.shape
        out = torch.empty(B, rec.T, C, dtype=torch.float32, device=x.device)
        with profiler.span('segmean_fwd', B=B, E_in=E_in, T=rec.T, C=C):
            _lib.check(_lib.lib().mcnn_segmean_fwd(x.data_ptr(), rec.rowptr.data_ptr(), rec.cols.data_ptr(), ec_out.data_ptr(),
                                                   out.data_ptr(), B, E_in, rec.T, C, rec.nnz_cap, _lib.current_stream()),
                       'mcnn_segmean_fwd')
        ctx.rec, ctx.ec_in, ctx.shape = rec, ec_in, (B, E_in, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        rec = ctx.rec
        B, E_in, C = ctx.shape
        dout = dout.contiguous()
        dx = torch.empty(B, E_in, C, dtype=torch.float32, device=dout.device)
        with profiler.span('segmean_bwd', B=B, E_in=E_in, T=rec.T, C=C):
            _lib.check(_lib.lib().mcnn_segmean_bwd(dout.data_ptr(), rec.rowptr.data_ptr(), rec.colptr.data_ptr(),
                                                   rec.rows.data_ptr(), ctx.ec_in.data_ptr(), dx.data_ptr(), B, E_in, rec.T, C,
                                                   rec.nnz_cap, _lib.current_stream()), 'mcnn_segmean_bwd')
        return dx, None, None, None


class UnpoolFn(torch.autograd.Function):
    """MeshUnpool.forward's matmul with the saved groups / occurrences (mesh_unpool.py:30-41)."""

    @staticmethod
    def forward(ctx, f, rec, ec_hi, ec_lo, E_hi):
        _chk_cuda(f)
        f = f.contiguous()
        B, E_lo, C = f.shape
        out = torch.empty(B, E_hi, C, dtype=torch.float32, device=f.device)
        with profiler.span('unpool_fwd', B=B, E_lo=E_lo, E_hi=E_hi, C=C):
            _lib.check(_lib.lib().mcnn_unpool_fwd(f.data_ptr(), rec.colptr.data_ptr(), rec.rows.data_ptr(), rec.occ.data_ptr(),
                                                  ec_hi.data_ptr(), out.data_ptr(), B, E_lo, E_hi, rec.E_in, C, rec.nnz_cap,
                                                  _lib.current_stream()), 'mcnn_unpool_fwd')
        ctx.rec, ctx.ec_lo, ctx.shape, ctx.E_hi = rec, ec_lo, (B, E_lo, C), E_hi
        return out

    @staticmethod
    def backward(ctx, dout):
        rec = ctx.rec
        B, E_lo, C = ctx.shape
        dout = dout.contiguous()
        df = torch.empty(B, E_lo, C, dtype=torch.float32, device=dout.device)
        with profiler.span('unpool_bwd', B=B, E_lo=E_lo, E_hi=ctx.E_hi, C=C):
            _lib.check(_lib.lib().mcnn_unpool_bwd(dout.data_ptr(), rec.rowptr.data_ptr(), rec.cols.data_ptr(), rec.occ.data_ptr(),
                                                  ctx.ec_lo.data_ptr(), df.data_ptr(), B, E_lo, ctx.E_hi, rec.E_in, C,
                                                  rec.nnz_cap, _lib.current_stream()), 'mcnn_unpool_bwd')
        return df, None, None, None, None


class FusedNormFn(torch.autograd.Function):
    """y = post(gamma * (pre(x [+ a]) - mean) * rstd + beta [+ r]) with BatchNorm / GroupNorm / InstanceNorm statistics
    (mcnn_norm_fwd / mcnn_norm_bwd); x, a, r: edge-major [B, E, C]."""

    @staticmethod
    def forward(ctx, x, a, r, gamma, beta, cfg):
        """cfg['pass_input']: returns (y, x) -- the caller uses the second output wherever it needs x again (the residual of
        MResConv, networks.py:171-179), so that both gradients of x arrive in this node's backward and the apply kernel adds
        them (mcnn_norm_bwd_acc) instead of a separate torch add."""
        _chk_cuda(x, a, r, gamma, beta)
        x_in = x
        x = x.contiguous()
        a = None if a is None else a.contiguous()
        r = None if r is None else r.contiguous()
        B, E, C = x.shape
        nblocks, rows, G = cfg['nblocks'], cfg['rows'], cfg['G']
        L = _lib.lib()
        dev = x.device
        y = torch.empty_like(x)
        stats_given = cfg.get('mean') is not None
        if stats_given:
            mean, rstd = cfg['mean'].contiguous(), cfg['rstd'].contiguous()
        else:
            mean = torch.empty(nblocks * G, dtype=torch.float32, device=dev)
            rstd = torch.empty(nblocks * G, dtype=torch.float32, device=dev)
        part = torch.empty(max(1, L.mcnn_norm_ws(nblocks, rows, C, G)), dtype=torch.float32, device=dev)
        rm, rv = cfg.get('running_mean'), cfg.get('running_var')
        with profiler.span('norm_fwd', nblocks=nblocks, rows=rows, C=C, G=G):
            _lib.check(L.mcnn_norm_fwd(x.data_ptr(), _lib.ptr(a), _lib.ptr(r), _lib.ptr(gamma), _lib.ptr(beta), y.data_ptr(),
                                       mean.data_ptr(), rstd.data_ptr(), part.data_ptr(), _lib.ptr(rm), _lib.ptr(rv),
                                       float(cfg.get('momentum', 0.1)), float(cfg['eps']), nblocks, rows, C, G,
                                       int(cfg['pre_relu']), int(cfg['post_relu']), int(stats_given), _lib.current_stream()),
                       'mcnn_norm_fwd')
        ctx.save_for_backward(x, a, gamma, mean, rstd, y)
        ctx.params = (gamma, beta)
        ctx.cfg = (nblocks, rows, C, G, int(cfg['pre_relu']), int(cfg['post_relu']), stats_given, r is not None)
        if cfg.get('pass_input'):
            return y, x_in
        return y

    @staticmethod
    def backward(ctx, dy, dpass=None):
        x, a, gamma, mean, rstd, y = ctx.saved_tensors
        nblocks, rows, C, G, pre_relu, post_relu, stats_given, has_r = ctx.cfg
        if stats_given:
            raise NotImplementedError('backward through a norm layer running on fixed (eval-mode) statistics')
        dy = dy.contiguous()
        L = _lib.lib()
        dev = dy.device
        dx = torch.empty_like(x)
        dr = torch.empty_like(x) if has_r else None
        direct = _direct_grad(*ctx.params) if (gamma is not None and ctx.needs_input_grad[3] and ctx.needs_input_grad[4]) else None
        if direct is not None:
            dgamma, dbeta = direct
        else:
            dgamma = torch.empty_like(gamma) if gamma is not None else None
            dbeta = torch.empty_like(gamma) if gamma is not None else None
        part = torch.empty(max(1, L.mcnn_norm_ws(nblocks, rows, C, G)), dtype=torch.float32, device=dev)
        s12 = torch.empty(2, nblocks * G, dtype=torch.float32, device=dev)
        # the gradient that reached x through the pass-through output is added inside the apply kernel -- unless x + a share dx
        dacc = None
        if dpass is not None and a is None and dpass.dtype == torch.float32 and dpass.shape == x.shape:
            dacc = dpass.contiguous()
        with profiler.span('norm_bwd', nblocks=nblocks, rows=rows, C=C, G=G):
            _lib.check(L.mcnn_norm_bwd_acc(dy.data_ptr(), y.data_ptr(), x.data_ptr(), _lib.ptr(a), _lib.ptr(gamma), mean.data_ptr(),
                                           rstd.data_ptr(), _lib.ptr(dacc), dx.data_ptr(), _lib.ptr(dr), _lib.ptr(dgamma),
                                           _lib.ptr(dbeta), part.data_ptr(), s12[0].data_ptr(), s12[1].data_ptr(), nblocks, rows,
                                           C, G, pre_relu, post_relu, _lib.current_stream()), 'mcnn_norm_bwd')
        if direct is not None:
            dgamma = dbeta = None                            # already in gamma.grad / beta.grad
        da = dx if a is not None else None
        if dpass is not None and dacc is None:
            dx = dx + dpass
        return dx, da, dr, dgamma, dbeta, None


class ClassifierHeadFn(torch.autograd.Function):
    """AvgPool1d over the edge axis -> fc1 -> relu -> fc2 (MeshConvNet.forward, models/networks.py:152-156) on the edge-major
    activation [B, E, C]: mcnn_head_fwd / mcnn_head_bwd, three launches per step instead of ~17 torch kernels."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _chk_cuda(x, w1, b1, w2, b2)
        x = x.contiguous()
        B, E, C = x.shape
        H, K = w1.shape[0], w2.shape[0]
        dev = x.device
        pooled = torch.empty(B, C, dtype=torch.float32, device=dev)
        h = torch.empty(B, H, dtype=torch.float32, device=dev)
        logits = torch.empty(B, K, dtype=torch.float32, device=dev)
        with profiler.span('head_fwd', B=B, E=E, C=C, H=H, K=K):
            _lib.check(_lib.lib().mcnn_head_fwd(x.data_ptr(), w1.data_ptr(), _lib.ptr(b1), w2.data_ptr(), _lib.ptr(b2),
                                                pooled.data_ptr(), h.data_ptr(), logits.data_ptr(), B, E, C, H, K,
                                                _lib.current_stream()), 'mcnn_head_fwd')
        ctx.save_for_backward(pooled, h, w1, w2)
        ctx.params = (w1, b1, w2, b2)
        ctx.dims = (B, E, C, H, K)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        pooled, h, w1, w2 = ctx.saved_tensors
        B, E, C, H, K = ctx.dims
        dev = dlogits.device
        dlogits = dlogits.contiguous()
        _, b1, _, b2 = ctx.params
        direct = _direct_grad(*ctx.params) if all(ctx.needs_input_grad[i] for i, p in enumerate(ctx.params, 1) if p is not None) else None
        if direct is not None:
            dw1, db1, dw2, db2 = direct
        else:
            dw1, dw2 = torch.empty_like(w1), torch.empty_like(w2)
            db1 = torch.empty_like(b1) if b1 is not None else None
            db2 = torch.empty_like(b2) if b2 is not None else None
        dx = torch.empty(B, E, C, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dh = torch.empty(B, H, dtype=torch.float32, device=dev)
        with profiler.span('head_bwd', B=B, E=E, C=C, H=H, K=K):
            _lib.check(_lib.lib().mcnn_head_bwd(dlogits.data_ptr(), pooled.data_ptr(), h.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                                dh.data_ptr(), _lib.ptr(dx), dw1.data_ptr(), _lib.ptr(db1), dw2.data_ptr(),
                                                _lib.ptr(db2), B, E, C, H, K, _lib.current_stream()), 'mcnn_head_bwd')
        if direct is not None:
            return dx, None, None, None, None
        return dx, dw1, db1, dw2, db2


# Cross-rank batch statistics for BatchNorm2d (MResConv.bn, networks.py:167; SURVEY 8e "optional SyncBN"): off by default, like
# any DDP run without SyncBN.  ddp.enable_sync_batchnorm() / MESHCNN_B200_SYNC_BN=1 switch it on.
SYNC_BN = os.environ.get('MESHCNN_B200_SYNC_BN', '0') == '1'
SYNC_BN_WEIGHT = None        # this rank's share of the global batch when the shards are uneven (None: equal shards)


class SyncBatchNormFn(torch.autograd.Function):
    """BatchNorm2d (training) with statistics over the rows of ALL ranks: the per-slab partial sums of the local rows
    (mcnn_norm_stats) are reduced here, all-reduced (one small NCCL call: 2C + 1 doubles), and handed back to the apply kernel
    (mcnn_norm_fwd with stats_given); the backward does the same with (sum dyh, sum dyh * xhat) between mcnn_norm_bwd_stats and
    mcnn_norm_bwd_apply.  dgamma / dbeta stay local sums: the flat gradient all-reduce averages them like every other parameter."""

    @staticmethod
    def forward(ctx, x, a, r, gamma, beta, cfg):
        import torch.distributed as dist
        _chk_cuda(x, a, r, gamma, beta)
        x = x.contiguous()
        a = None if a is None else a.contiguous()
        r = None if r is None else r.contiguous()
        B, E, C = x.shape
        rows = B * E
        L = _lib.lib()
        st = _lib.current_stream()
        slabs = L.mcnn_norm_slabs(1, rows, C, C)
        part = torch.empty(slabs * 2 * C, dtype=torch.float32, device=x.device)
        _lib.check(L.mcnn_norm_stats(x.data_ptr(), _lib.ptr(a), part.data_ptr(), 1, rows, C, C, int(cfg['pre_relu']), st), 'mcnn_norm_stats')
        buf = torch.cat([part.view(slabs, 2, C).double().sum(0).reshape(-1), torch.full((1,), float(rows), dtype=torch.float64, device=x.device)])
        dist.all_reduce(buf)
        n = buf[2 * C]
        mean = buf[:C] / n
        var = (buf[C:2 * C] / n - mean * mean).clamp_(min=0.0)
        rstd = torch.rsqrt(var + float(cfg['eps']))
        rm, rv = cfg.get('running_mean'), cfg.get('running_var')
        if rm is not None:
            mom = float(cfg.get('momentum', 0.1))
            with torch.no_grad():
                rm.mul_(1.0 - mom).add_(mean.to(rm.dtype), alpha=mom)
                rv.mul_(1.0 - mom).add_((var * n / (n - 1.0).clamp(min=1.0)).to(rv.dtype), alpha=mom)
        mean32, rstd32 = mean.float().contiguous(), rstd.float().contiguous()
        y = torch.empty_like(x)
        _lib.check(L.mcnn_norm_fwd(x.data_ptr(), _lib.ptr(a), _lib.ptr(r), _lib.ptr(gamma), _lib.ptr(beta), y.data_ptr(),
                                   mean32.data_ptr(), rstd32.data_ptr(), part.data_ptr(), None, None, 0.0, float(cfg['eps']), 1, rows, C, C,
                                   int(cfg['pre_relu']), int(cfg['post_relu']), 1, st), 'mcnn_norm_fwd')
        ctx.save_for_backward(x, a, gamma, mean32, rstd32, y, n)
        ctx.cfg = (rows, C, slabs, int(cfg['pre_relu']), int(cfg['post_relu']), r is not None)
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        x, a, gamma, mean, rstd, y, n = ctx.saved_tensors
        rows, C, slabs, pre_relu, post_relu, has_r = ctx.cfg
        dy = dy.contiguous()
        L = _lib.lib()
        st = _lib.current_stream()
        part = torch.empty(slabs * 2 * C, dtype=torch.float32, device=dy.device)
        _lib.check(L.mcnn_norm_bwd_stats(dy.data_ptr(), y.data_ptr(), x.data_ptr(), _lib.ptr(a), mean.data_ptr(), rstd.data_ptr(),
                                         part.data_ptr(), 1, rows, C, C, pre_relu, post_relu, st), 'mcnn_norm_bwd_stats')
        ab = part.view(slabs, 2, C).double().sum(0)                      # local sum dyh, sum dyh * xhat per channel
        dbeta, dgamma = ab[0].float(), ab[1].float()
        w = SYNC_BN_WEIGHT
        tot = ab.clone() if w is None else ab * float(w)                 # uneven shards: the global loss weighs rank r by its share
        dist.all_reduce(tot)
        if w is not None:
            tot = tot / float(w)
        gm = gamma.double() if gamma is not None else torch.ones(C, dtype=torch.float64, device=dy.device)
        s1 = (gm * tot[0] / n).float().contiguous()
        s2 = (gm * tot[1] / n).float().contiguous()
        dx = torch.empty_like(x)
        dr = torch.empty_like(x) if has_r else None
        _lib.check(L.mcnn_norm_bwd_apply(dy.data_ptr(), y.data_ptr(), x.data_ptr(), _lib.ptr(a), _lib.ptr(gamma), mean.data_ptr(),
                                         rstd.data_ptr(), s1.data_ptr(), s2.data_ptr(), dx.data_ptr(), _lib.ptr(dr), 1, rows, C, C,
                                         pre_relu, post_relu, st), 'mcnn_norm_bwd_apply')
        if gamma is None:
            dgamma = dbeta = None
        return dx, (dx if a is not None else None), dr, dgamma, dbeta, None


def _sync_bn_active():
    if not SYNC_BN:
        return False
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def fused_norm(x, norm, pre_add=None, pre_relu=False, post_add=None, post_relu=False, pass_input=False):
    """Apply a torch norm module (BatchNorm2d / GroupNorm / InstanceNorm2d, the three the reference's networks use) to
    edge-major x [B, E, C] with the surrounding add / relu fused in.  Returns None if `norm` is not one of those."""
    import torch.nn as nn
    B, E, C = x.shape
    if C % 4 != 0:
        return None
    cfg = {'pre_relu': pre_relu, 'post_relu': post_relu, 'eps': norm.eps if hasattr(norm, 'eps') else 1e-5}
    gamma = beta = None
    if isinstance(norm, nn.GroupNorm):
        cfg.update(nblocks=B, rows=E, G=norm.num_groups)
        gamma, beta = norm.weight, norm.bias
    elif isinstance(norm, nn.InstanceNorm2d):
        if norm.track_running_stats:
            return None
        cfg.update(nblocks=B, rows=E, G=C)
        gamma, beta = norm.weight, norm.bias
    elif isinstance(norm, nn.BatchNorm2d):
        cfg.update(nblocks=1, rows=B * E, G=C)
        gamma, beta = norm.weight, norm.bias
        if norm.training or not norm.track_running_stats:
            if norm.track_running_stats:
                if norm.momentum is None:
                    return None
                cfg.update(running_mean=norm.running_mean, running_var=norm.running_var, momentum=norm.momentum)
                norm.num_batches_tracked.add_(1)
            if _sync_bn_active():
                y = SyncBatchNormFn.apply(x, pre_add, post_add, gamma, beta, cfg)
                return (y, x) if pass_input else y
        else:
            cfg.update(mean=norm.running_mean, rstd=torch.rsqrt(norm.running_var + norm.eps))
    else:
        return None
    if (gamma is None) != (beta is None):
        return None
    if pass_input:
        cfg['pass_input'] = True
    return FusedNormFn.apply(x, pre_add, post_add, gamma, beta, cfg)
