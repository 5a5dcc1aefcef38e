"""Layout plumbing between the reference's logical [B, C, E(,1)] tensors and the edge-major [B, E, C] memory the
kernels use.  A [B, C, E, 1] tensor in torch.channels_last memory format *is* edge-major memory, so conversions are
free views whenever the producer was one of our own layers."""
import torch


def to_edge_major(x):
    """[B, C, E] or [B, C, E, 1] (any strides) -> contiguous [B, E, C] fp32, without a copy when possible."""
    if x.dim() == 4:
        x = x.squeeze(-1)
    if x.dim() != 3:
        raise ValueError('expected [B, C, E] or [B, C, E, 1], got %s' % (tuple(x.shape),))
    if x.dtype != torch.float32:
        raise TypeError('meshcnn_b200 kernels compute in fp32; got %s' % x.dtype)
    v = x.permute(0, 2, 1)
    return v if v.is_contiguous() else v.contiguous()


def as_bce1(y):
    """edge-major [B, E, C] -> logical [B, C, E, 1] view (channels_last strides)."""
    return y.permute(0, 2, 1).unsqueeze(-1)


def as_bce(y):
    """edge-major [B, E, C] -> logical [B, C, E] view."""
    return y.permute(0, 2, 1)
