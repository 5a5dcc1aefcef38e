"""Drop-in for models/layers/mesh_conv.py: same constructor, same ``conv`` sub-module (so checkpoints load,
SURVEY F17), same call signature and output shape -- the work is one fused CUDA kernel instead of
pad_gemm + create_GeMM + Conv2d."""
import torch
import torch.nn as nn

from .functional import MeshConvFn
from .layout import as_bce1, to_edge_major
from .mesh import MeshBatch


class MeshConv(nn.Module):
    """Convolution between an edge and its 4 one-ring neighbours (mesh_conv.py:5-26).
    x: [B, Cin, E] or [B, Cin, E, 1];  mesh: sequence of B Mesh objects;  returns [B, Cout, E, 1]."""

    def __init__(self, in_channels, out_channels, k=5, bias=True):
        super(MeshConv, self).__init__()
        if k != 5:
            raise NotImplementedError('the one-ring has 4 neighbours: k must be 5 (mesh_conv.py:12-15)')
        # parameter container only (weight [Cout, Cin, 1, 5], bias [Cout]); its forward is never called
        self.conv = nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=(1, k), bias=bias)
        self.k = k
        self._wt_ws = None
        self._img = None            # (fwd image, bwd image, event, step token) prepared ahead of time by prepare_images()

    def __call__(self, edge_f, mesh):
        return self.forward(edge_f, mesh)

    def forward(self, x, mesh):
        return as_bce1(self.forward_em(to_edge_major(x), MeshBatch.of(mesh, x.shape[2], x.device)))

    def prepare_images(self, token, stream):
        """Build the tensor-core weight images of this layer on `stream` (a side stream at the start of a step: the weight
        does not depend on the activations).  forward / backward of step `token` then skip their own re-layout kernels."""
        from . import _lib
        from . import functional as Fn
        w = self.conv.weight
        cout, cin = w.shape[0], w.shape[1]
        L = _lib.lib()
        if not (Fn.USE_TC and w.is_cuda and L.mcnn_meshconv_tc_supported(cin, cout)):
            return
        if self._img is None or self._img[0].device != w.device:
            n = 2 * 5 * cin * cout
            self._img = [torch.empty(n, dtype=torch.float32, device=w.device), torch.empty(n, dtype=torch.float32, device=w.device),
                         torch.cuda.Event(), -1]
        with torch.cuda.stream(stream):
            _lib.check(L.mcnn_meshconv_tc_weight_images(w.data_ptr(), self._img[0].data_ptr(), self._img[1].data_ptr(), cin, cout,
                                                        stream.cuda_stream), 'mcnn_meshconv_tc_weight_images')
            self._img[2].record(stream)
        self._img[3] = token

    def forward_em(self, x_em, batch, pass_input=False):
        """Edge-major entry used by meshcnn_b200.networks: [B, E, Cin] -> [B, E, Cout]; pass_input=True: (out, x_em), see
        MeshConvFn.forward."""
        w = self.conv.weight
        if self._wt_ws is None or self._wt_ws.device != x_em.device:
            self._wt_ws = torch.empty(2 * w.shape[1] * 5, w.shape[0], dtype=torch.float32, device=x_em.device)
        from . import functional as Fn
        img = self._img if (self._img is not None and self._img[3] == Fn.STEP_TOKEN and Fn.STEP_TOKEN >= 0) else None
        if pass_input:
            return MeshConvFn.apply(x_em, w, self.conv.bias, batch.level(x_em.shape[1]), self._wt_ws, img, True)
        return MeshConvFn.apply(x_em, w, self.conv.bias, batch.level(x_em.shape[1]), self._wt_ws, img)
