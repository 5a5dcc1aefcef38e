"""MeshConv CUDA path (through the C ABI) vs the reference-generated golden vectors and vs the oracle.
Tolerance: 1e-4 relative, norm-wise (BASELINE.json north_star, fp32)."""
import numpy as np
import pytest
import torch

from tests.helpers import close, golden, golden_names, prepared

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.mark.parametrize('name', golden_names('conv_'))
def test_conv_golden(name):
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, cuda, product_mesh
    d = golden(name)
    B = int(d['B'])
    meshes = [product_mesh(d['M%d_vs0' % i], d['M%d_faces' % i]) for i in range(B)]
    cout, cin = d['w'].shape[:2]
    conv = MeshConv(cin, cout, bias='b' in d.files).to(DEV)
    with torch.no_grad():
        conv.conv.weight.copy_(cuda(d['w']))
        if 'b' in d.files:
            conv.conv.bias.copy_(cuda(d['b']))
    x = cuda(d['x']).requires_grad_(True)
    y = conv(x, meshes)
    assert tuple(y.shape) == d['y'].shape
    assert close(y.detach().cpu().numpy(), d['y'], RTOL)
    y.backward(cuda(d['dy']))
    assert close(x.grad.cpu().numpy(), d['dx'], RTOL)
    assert close(conv.conv.weight.grad.cpu().numpy(), d['dw'], RTOL)
    if 'b' in d.files:
        assert close(conv.conv.bias.grad.cpu().numpy(), d['db'], RTOL)


@pytest.mark.parametrize('cin,cout,bias', [(5, 64, False), (64, 64, False), (128, 256, True), (32, 8, True), (7, 9, True),
                                           (256, 256, False), (64, 128, True), (32, 64, True)])
def test_conv_vs_oracle_random(cin, cout, bias):
    """Config-sized layers on a mixed batch (closed meshes, one open mesh, short meshes -> padded slots)."""
    from oracle import meshcnn_oracle as O
    from meshcnn_b200 import synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, cuda, product_mesh
    specs = [('ico', dict(nu=5), 0), ('torus', dict(m=10, n=25), 1), ('grid', dict(m=12, n=14), 2), ('ico', dict(nu=4), 3)]
    E = 750
    rs = np.random.RandomState(cin * 1000 + cout)
    prods, oracles = [], []
    for kind, kw, seed in specs:
        vs, faces = synth.make(kind, seed, **kw)
        prods.append(product_mesh(vs, faces))
        oracles.append(O.OracleMesh(prepared(vs, faces)))
    B = len(prods)
    xh = rs.randn(B, cin, E).astype(np.float32)
    wh = (rs.randn(cout, cin, 1, 5) / np.sqrt(5 * cin)).astype(np.float32)
    bh = rs.randn(cout).astype(np.float32) if bias else None
    dyh = rs.randn(B, cout, E, 1).astype(np.float32)
    xo = torch.from_numpy(xh).requires_grad_(True)
    wo = torch.from_numpy(wh).requires_grad_(True)
    bo = torch.from_numpy(bh).requires_grad_(True) if bias else None
    yo = O.conv_forward(xo, oracles, wo, bo)
    yo.backward(torch.from_numpy(dyh))
    conv = MeshConv(cin, cout, bias=bias).to(DEV)
    with torch.no_grad():
        conv.conv.weight.copy_(cuda(wh))
        if bias:
            conv.conv.bias.copy_(cuda(bh))
    x = cuda(xh).requires_grad_(True)
    y = conv(x, prods)
    y.backward(cuda(dyh))
    assert close(y.detach().cpu().numpy(), yo.detach().numpy(), RTOL)
    assert close(x.grad.cpu().numpy(), xo.grad.numpy(), RTOL)
    assert close(conv.conv.weight.grad.cpu().numpy(), wo.grad.numpy(), RTOL)
    if bias:
        assert close(conv.conv.bias.grad.cpu().numpy(), bo.grad.numpy(), RTOL)
    # gradients on padded slots are exactly zero
    for b, m in enumerate(oracles):
        assert float(x.grad[b, :, m.edges_count:].abs().max().item() if m.edges_count < E else 0.0) == 0.0


def test_conv_linearity_full_size():
    """Size-independent property at the bench configuration (B=64, 750 edges): conv(a x1 + x2) == a conv(x1) + conv(x2)
    does NOT hold (|.| is not linear) but conv is positively homogeneous: conv(a x) = a conv(x) for a > 0, bias-free."""
    from meshcnn_b200 import synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico', s, nu=5)) for s in range(64)]
    conv = MeshConv(64, 128, bias=False).to(DEV)
    x = torch.randn(64, 64, 750, device=DEV)
    y1 = conv(x, meshes)
    y2 = conv(2.5 * x, meshes)
    assert torch.allclose(2.5 * y1, y2, rtol=1e-5, atol=1e-5)
    assert y1.shape == (64, 128, 750, 1)


@pytest.mark.parametrize('cin,cout', [(64, 64), (128, 128), (256, 256), (128, 256)])
def test_conv_tensor_core_path_matches_ffma_path(cin, cout):
    """tcgen05 3xTF32 forward vs the exact-fp32 FFMA forward on a bench-sized batch: <= 3e-5 relative (both sides carry their own ~1e-6 rounding at K = 1280)."""
    from meshcnn_b200 import functional as Fn
    from meshcnn_b200 import synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    B = 24
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    assert Fn.USE_TC
    y_tc = conv(x, meshes)
    Fn.USE_TC = False
    try:
        y_ff = conv(x, meshes)
    finally:
        Fn.USE_TC = True
    assert close(y_tc.detach().cpu().numpy(), y_ff.detach().cpu().numpy(), 3e-5)
    # backward: tensor-core dgrad / wgrad vs the FFMA kernels
    dy = torch.randn_like(y_tc)
    xg = x.clone().requires_grad_(True)
    conv.zero_grad()
    conv(xg, meshes).backward(dy)
    g_tc = [xg.grad.clone(), conv.conv.weight.grad.clone(), conv.conv.bias.grad.clone()]
    Fn.USE_TC = False
    try:
        xg2 = x.clone().requires_grad_(True)
        conv.zero_grad()
        conv(xg2, meshes).backward(dy)
    finally:
        Fn.USE_TC = True
    g_ff = [xg2.grad, conv.conv.weight.grad, conv.conv.bias.grad]
    for a, b in zip(g_tc, g_ff):
        assert close(a.cpu().numpy(), b.cpu().numpy(), 3e-5)


@pytest.mark.parametrize('mode', [1, 2], ids=['cta_pair', 'producers16'])
@pytest.mark.parametrize('cin,cout,B', [(128, 256, 24), (256, 256, 23), (256, 512, 5), (128, 128, 7)])
def test_conv_cta_pair_kernel_matches_single_cta_kernel(cin, cout, B, mode):
    """The cta_group::2 forward kernel (two CTAs share the weight tile) against the single-CTA tcgen05 kernel: both run the
    same 3xTF32 products in the same K order, so the results must agree to rounding of the accumulation order (bit-exact
    in practice); odd tile counts exercise the surplus CTA of the last pair."""
    from meshcnn_b200 import _lib, synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    from meshcnn_b200 import functional as Fn
    L = _lib.lib()
    sk = Fn.USE_SK
    try:
        Fn.USE_SK = False                        # both are one-CTA(-pair)-per-tile kernels
        L.mcnn_debug_set_tc_pair(0)
        y1 = conv(x, meshes).clone()
        L.mcnn_debug_set_tc_pair(mode)           # 1: cta_group::2 kernel, 2: 16 producer warps
        y2 = conv(x, meshes).clone()
        torch.cuda.synchronize()
    finally:
        Fn.USE_SK = sk
        L.mcnn_debug_set_tc_pair(_lib.TC_PAIR_DEFAULT)
    assert close(y2.detach().cpu().numpy(), y1.detach().cpu().numpy(), 1e-6)


@pytest.mark.parametrize('ctas', [0, 1, 7, 148, 251], ids=lambda c: 'ctas%d' % c)
@pytest.mark.parametrize('cin,cout,B', [(64, 64, 24), (128, 256, 9), (256, 256, 23), (256, 512, 5), (32, 192, 3)])
def test_conv_stream_k_kernel_matches_tile_kernel(cin, cout, B, ctas):
    """The persistent stream-K forward kernel against the one-CTA-per-tile kernel, with the grid forced to 1 CTA (no tile is
    cut), 7 (long ranges), 148 and 251 (more CTAs than tiles for the small cases: a tile is cut into 3+ partials).  Same
    products; only the fp32 sum of the partials of a cut tile is associated differently.  The signs kept for the backward
    must be identical, and a second launch on the same workspace must reproduce the first bit for bit (flags lowered)."""
    from meshcnn_b200 import _lib, synth
    from meshcnn_b200 import functional as Fn
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)
    L = _lib.lib()
    assert Fn.USE_SK and Fn.USE_TC

    def run():
        xg = x.clone().requires_grad_(True)
        y = conv(xg, meshes)
        y.backward(dy)
        torch.cuda.synchronize()
        return y.detach().clone(), xg.grad.clone()
    try:
        Fn.USE_SK = False
        y1, g1 = run()
        Fn.USE_SK = True
        L.mcnn_debug_set_tc_sk_ctas(ctas)
        y2, g2 = run()
        y3, g3 = run()
    finally:
        Fn.USE_SK = True
        L.mcnn_debug_set_tc_sk_ctas(0)
    assert torch.equal(y2, y3)
    # measured: both differ from the exact-fp32 FFMA kernel by <= 1.1e-5 of max|y| (fp32 accumulation in TMEM over K = 1280);
    # summing a cut tile's partials moves the result by the same order
    assert close(y2.cpu().numpy(), y1.cpu().numpy(), 2e-5)
    assert close(g2.cpu().numpy(), g1.cpu().numpy(), 1e-5)          # same signs; the scatter atomics reorder sums
    if ctas == 1:
        assert torch.equal(y2, y1)


@pytest.mark.parametrize('cin,cout,B', [(64, 64, 24), (128, 256, 9), (256, 256, 23), (256, 64, 5), (32, 192, 3), (256, 256, 1)])
def test_conv_persistent_data_gradient_matches_tile_kernel(cin, cout, B):
    """The persistent data-gradient kernel (dedicated producer / MMA / scatter warps walking a range of items) against the
    one-CTA-per-item kernel: identical products and signs, only the order of the scatter atomics differs."""
    from meshcnn_b200 import _lib, synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)
    L = _lib.lib()

    def run():
        xg = x.clone().requires_grad_(True)
        conv(xg, meshes).backward(dy)
        torch.cuda.synchronize()
        return xg.grad.clone()
    try:
        L.mcnn_debug_set_tc_bd_tile(1)
        g1 = run()
        L.mcnn_debug_set_tc_bd_tile(0)
        g2 = run()
    finally:
        L.mcnn_debug_set_tc_bd_tile(0)
    assert close(g2.cpu().numpy(), g1.cpu().numpy(), 2e-6)


@pytest.mark.parametrize('cin,cout,B', [(128, 256, 9), (256, 256, 23), (64, 512, 5), (32, 256, 1)])
def test_conv_weight_gradient_two_output_tiles_per_cta(cin, cout, B):
    """Weight-gradient kernel with two 128-channel output tiles per CTA (gathered operand built once for 256 output
    channels) against one tile per CTA: same products, the row range is split over a different number of CTAs, so the
    results agree to fp32 rounding of the split sums."""
    from meshcnn_b200 import _lib, synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)
    L = _lib.lib()
    res = []
    try:
        for ot in (1, 2):
            L.mcnn_debug_set_tc_bw_ot(ot)
            conv.zero_grad()
            conv(x, meshes).backward(dy)
            torch.cuda.synchronize()
            res.append((conv.conv.weight.grad.clone(), conv.conv.bias.grad.clone()))
    finally:
        L.mcnn_debug_set_tc_bw_ot(0)
    for a, b in zip(res[1], res[0]):
        err = float((a - b).abs().max() / b.abs().max())
        assert err <= 1e-5, err


@pytest.mark.parametrize('cin,cout,B', [(128, 256, 9), (256, 256, 23), (64, 128, 5), (32, 64, 3), (64, 192, 2), (256, 256, 64)])
def test_conv_weight_gradient_tmem_operand_matches_smem_kernel(cin, cout, B):
    """Weight-gradient kernel with dout^T held in tensor memory (the default) against the first-generation kernel that keeps
    both operands in shared memory: the same split of the edge range, the same three TF32 products per term, so the results
    agree to the rounding of the accumulation order inside the tensor core (<= 1e-6); and both against exact fp32 on the
    CPU oracle formula at 1e-4."""
    from meshcnn_b200 import _lib, synth
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)
    L = _lib.lib()
    res = []
    try:
        for on in (0, 1):
            L.mcnn_debug_set_tc_bw_tmem_a(on)
            conv.zero_grad()
            conv(x, meshes).backward(dy)
            torch.cuda.synchronize()
            res.append((conv.conv.weight.grad.clone(), conv.conv.bias.grad.clone()))
    finally:
        L.mcnn_debug_set_tc_bw_tmem_a(1)
    for a, b in zip(res[1], res[0]):
        err = float((a - b).abs().max() / b.abs().max())
        assert err <= 2e-6, err


@pytest.mark.parametrize('cin,cout,B', [(256, 256, 26), (256, 256, 39), (128, 128, 52), (64, 64, 64)])
def test_conv_stream_k_at_bench_tile_counts(cin, cout, B):
    """Tile counts just above a multiple of the SM count, as in the bench configuration (153 / 229 / 305 / 375 tiles of
    128 edges on 148 SMs): every CTA holds cut pieces, most own one.  Forward and the gradients that use its signs against
    the one-CTA-per-tile kernel."""
    from meshcnn_b200 import synth
    from meshcnn_b200 import functional as Fn
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)

    def run():
        xg = x.clone().requires_grad_(True)
        y = conv(xg, meshes)
        y.backward(dy)
        torch.cuda.synchronize()
        return y.detach().clone(), xg.grad.clone()
    try:
        Fn.USE_SK = False
        y1, g1 = run()
        Fn.USE_SK = True
        y2, g2 = run()
    finally:
        Fn.USE_SK = True
    assert close(y2.cpu().numpy(), y1.cpu().numpy(), 2e-5)
    assert close(g2.cpu().numpy(), g1.cpu().numpy(), 1e-5)


@pytest.mark.parametrize('cin,cout,B', [(64, 64, 8), (128, 256, 9), (256, 256, 20)])
def test_conv_single_pass_tf32_mode_is_within_the_reduced_precision_tolerance(cin, cout, B):
    """Opt-in reduced precision (meshcnn_b200.set_precision('tf32'): ONE TF32 MMA per product instead of three): forward,
    data gradient and weight gradient against the default fp32-accurate kernels.  BASELINE.json allows a reduced-precision
    path 2e-2 relative; TF32 operands (10 mantissa bits) with fp32 accumulation land near 5e-4.  The lower bound proves the
    switch really changes the arithmetic."""
    from meshcnn_b200 import synth
    from meshcnn_b200 import functional as Fn
    from meshcnn_b200.mesh_conv import MeshConv
    from tests.gpu_util import DEV, product_mesh
    meshes = [product_mesh(*synth.make('ico' if s % 3 else 'torus', s, **(dict(nu=5) if s % 3 else dict(m=10, n=25)))) for s in range(B)]
    torch.manual_seed(cin + cout + B)
    conv = MeshConv(cin, cout, bias=True).to(DEV)
    x = torch.randn(B, cin, 750, device=DEV)
    dy = torch.randn(B, cout, 750, 1, device=DEV)

    def run():
        conv.zero_grad()
        xg = x.clone().requires_grad_(True)
        y = conv(xg, meshes)
        y.backward(dy)
        torch.cuda.synchronize()
        return [t.detach().clone() for t in (y, xg.grad, conv.conv.weight.grad, conv.conv.bias.grad)]
    try:
        ref = run()
        Fn.set_precision('tf32')
        got = run()
    finally:
        Fn.set_precision('fp32')
    errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, ref)]
    assert all(e <= 2e-2 for e in errs), errs
    assert all(e <= 5e-3 for e in errs), errs                  # what single-pass TF32 actually delivers
    assert min(errs[:3]) > 2e-6, errs                          # the mode is really on
