"""Network assemblies with the call signatures and parameter names of the reference's models/networks.py, so that
``latest_net.pth`` files interchange and train.py-style code can build them through
``define_classifier``.  The reference file itself also works unchanged on top of the drop-in layers
(meshcnn_b200.dropin); this module exists because the GPU box has no reference tree, and because running the
blocks directly on edge-major tensors removes every layout conversion between layers.

Module tree (== state-dict keys) mirrored from the reference:
  MeshConvNet          conv{i} (MResConv: conv0, bn{j}, conv{j}), norm{i}, pool{i}, gp, fc1, fc2      networks.py:124-157
  MeshEncoderDecoder   encoder.convs[i] (DownConv: conv1, conv2[], bn[], pool)                           networks.py:182-346
                       decoder.up_convs[i] / decoder.final_conv (UpConv: up_conv, conv1, conv2[], bn[], unroll)

Internally activations are edge-major [B, E, C]; normalisation layers act on the channels_last [B, C, E, 1]
view of the same memory, so their statistics include padded edge slots exactly like the reference's.
"""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init
from torch.optim import lr_scheduler

from .functional import ClassifierHeadFn, fused_norm
from .layout import as_bce, as_bce1, to_edge_major
from .mesh import MeshBatch
from .mesh_conv import MeshConv

import os
# residual branches hand their gradient to the conv's data-gradient kernel as its starting value (MeshConvFn pass_input);
# MESHCNN_B200_RES_PASS=0: plain autograd accumulation (A/B switch)
RES_PASS = os.environ.get('MESHCNN_B200_RES_PASS', '1') != '0'
from .mesh_pool import MeshPool
from .mesh_unpool import MeshUnpool


# ---------------------------------------------------------------------------------------------------------------
# factories (networks.py:17-118)
# ---------------------------------------------------------------------------------------------------------------
class NoNorm(nn.Module):
    def __init__(self, fake=True):
        super(NoNorm, self).__init__()
        self.fake = fake

    def forward(self, x):
        return x


def get_norm_layer(norm_type='instance', num_groups=1):
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type == 'group':
        return functools.partial(nn.GroupNorm, affine=True, num_groups=num_groups)
    if norm_type == 'none':
        return NoNorm
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_norm_args(norm_layer, nfeats_list):
    """The reference only handles 'group' and 'none' here; 'batch' / 'instance' are accepted too."""
    if norm_layer is NoNorm:
        return [{'fake': True} for _ in nfeats_list]
    name = norm_layer.func.__name__
    if name == 'GroupNorm':
        return [{'num_channels': f} for f in nfeats_list]
    if name in ('BatchNorm2d', 'InstanceNorm2d'):
        return [{'num_features': f} for f in nfeats_list]
    raise NotImplementedError('normalization layer [%s] is not found' % name)


def get_scheduler(optimizer, opt):
    if opt.lr_policy == 'lambda':
        def rule(epoch):
            return 1.0 - max(0, epoch + 1 + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=rule)
    if opt.lr_policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    raise NotImplementedError('learning rate policy [%s] is not implemented' % opt.lr_policy)


def init_weights(net, init_type, init_gain):
    """networks.py:65-82: every module whose class name contains Conv / Linear and owns a ``weight``."""
    fns = {'normal': lambda w: init.normal_(w, 0.0, init_gain),
           'xavier': lambda w: init.xavier_normal_(w, gain=init_gain),
           'kaiming': lambda w: init.kaiming_normal_(w, a=0, mode='fan_in'),
           'orthogonal': lambda w: init.orthogonal_(w, gain=init_gain)}
    if init_type not in fns:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)

    def visit(m):
        cls = m.__class__.__name__
        if hasattr(m, 'weight') and ('Conv' in cls or 'Linear' in cls):
            fns[init_type](m.weight.data)
        elif 'BatchNorm2d' in cls:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    net.apply(visit)


def init_net(net, init_type, init_gain, gpu_ids):
    """networks.py:85-93.  One process drives one GPU; multi-GPU training goes through meshcnn_b200.ddp (the
    reference's nn.DataParallel replicates the mesh list to every replica and cannot work)."""
    if len(gpu_ids) > 1:
        raise NotImplementedError('use one process per GPU (meshcnn_b200.ddp / torchrun) instead of gpu_ids=%s' % (gpu_ids,))
    if len(gpu_ids) == 1:
        assert torch.cuda.is_available()
        net.cuda(gpu_ids[0])
        net = torch.nn.DataParallel(net, gpu_ids)      # single-device wrapper: keeps `.module` for save_network
    if init_type != 'none':
        init_weights(net, init_type, init_gain)
    return net


def define_classifier(input_nc, ncf, ninput_edges, nclasses, opt, gpu_ids, arch, init_type, init_gain):
    norm_layer = get_norm_layer(norm_type=opt.norm, num_groups=opt.num_groups)
    if arch == 'mconvnet':
        net = MeshConvNet(norm_layer, input_nc, ncf, nclasses, ninput_edges, opt.pool_res, opt.fc_n, opt.resblocks)
    elif arch == 'meshunet':
        net = MeshEncoderDecoder([ninput_edges] + opt.pool_res, [input_nc] + ncf, ncf[::-1] + [nclasses],
                                 blocks=opt.resblocks, transfer_data=True)
    else:
        raise NotImplementedError('Encoder model name [%s] is not recognized' % arch)
    return init_net(net, init_type, init_gain, gpu_ids)


def define_loss(opt):
    if opt.dataset_mode == 'classification':
        return torch.nn.CrossEntropyLoss()
    if opt.dataset_mode == 'segmentation':
        return torch.nn.CrossEntropyLoss(ignore_index=-1)
    raise NotImplementedError(opt.dataset_mode)


# ---------------------------------------------------------------------------------------------------------------
# edge-major helpers
# ---------------------------------------------------------------------------------------------------------------
def _norm_em(norm, x, pre_add=None, pre_relu=False, post_add=None, post_relu=False, pass_input=False):
    """post(norm(pre(x [+ pre_add])) [+ post_add]) on edge-major data.  BatchNorm2d / GroupNorm / InstanceNorm2d run as one
    fused CUDA block (mcnn_norm_fwd); anything else goes through torch on the [B, C, E, 1] channels_last view."""
    y = fused_norm(x, norm, pre_add, pre_relu, post_add, post_relu, pass_input)     # pass_input: (y, x), see FusedNormFn.forward
    if y is not None:
        return y
    z = x if pre_add is None else x + pre_add
    if pre_relu:
        z = F.relu(z)
    y = to_edge_major(norm(as_bce1(z)))
    if post_add is not None:
        y = y + post_add
    y = F.relu(y) if post_relu else y
    return (y, x) if pass_input else y


# ---------------------------------------------------------------------------------------------------------------
# classification (networks.py:124-179)
# ---------------------------------------------------------------------------------------------------------------
class MResConv(nn.Module):
    def __init__(self, in_channels, out_channels, skips=1):
        super(MResConv, self).__init__()
        self.in_channels, self.out_channels, self.skips = in_channels, out_channels, skips
        self.conv0 = MeshConv(in_channels, out_channels, bias=False)
        for i in range(1, skips + 1):
            self.add_module('bn%d' % i, nn.BatchNorm2d(out_channels))
            self.add_module('conv%d' % i, MeshConv(out_channels, out_channels, bias=False))

    def forward_em(self, x, batch, defer_tail=False):
        """defer_tail=True returns (x, skip) and leaves the final relu(x + skip) to the caller's fused norm block."""
        x = self.conv0.forward_em(x, batch)
        skip = x
        for i in range(1, self.skips + 1):
            if i == 1 and RES_PASS:          # skip is conv0's output again: both of its gradients meet in bn1's backward kernel
                x, skip = _norm_em(getattr(self, 'bn%d' % i), x, pre_relu=True, pass_input=True)
            else:
                x = _norm_em(getattr(self, 'bn%d' % i), x, pre_relu=True)
            x = getattr(self, 'conv%d' % i).forward_em(x, batch)
        if defer_tail:
            return x, skip
        return F.relu(x + skip)

    def forward(self, x, mesh):
        return as_bce1(self.forward_em(to_edge_major(x), MeshBatch.of(mesh, x.shape[2], x.device)))


class MeshConvNet(nn.Module):
    """Network for learning a global shape descriptor (classification)."""

    def __init__(self, norm_layer, nf0, conv_res, nclasses, input_res, pool_res, fc_n, nresblocks=3):
        super(MeshConvNet, self).__init__()
        self.k = [nf0] + conv_res
        self.res = [input_res] + pool_res
        norm_args = get_norm_args(norm_layer, self.k[1:])
        for i in range(len(self.k) - 1):
            self.add_module('conv%d' % i, MResConv(self.k[i], self.k[i + 1], nresblocks))
            self.add_module('norm%d' % i, norm_layer(**norm_args[i]))
            self.add_module('pool%d' % i, MeshPool(self.res[i + 1]))
        self.gp = torch.nn.AvgPool1d(self.res[-1])
        self.fc1 = nn.Linear(self.k[-1], fc_n)
        self.fc2 = nn.Linear(fc_n, nclasses)

    def forward(self, x, mesh):
        batch = MeshBatch.of(mesh, x.shape[2], x.device)
        x = to_edge_major(x)
        for i in range(len(self.k) - 1):
            x, skip = getattr(self, 'conv%d' % i).forward_em(x, batch, defer_tail=True)
            x = _norm_em(getattr(self, 'norm%d' % i), x, pre_add=skip, pre_relu=True, post_relu=True)
            x = getattr(self, 'pool%d' % i).forward_em(x, batch)
        # AvgPool1d(res[-1]) over the whole (zero-padded) edge axis, fc1 + relu, fc2: one fused kernel
        return ClassifierHeadFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


# ---------------------------------------------------------------------------------------------------------------
# segmentation (networks.py:182-388)
# ---------------------------------------------------------------------------------------------------------------
def _xavier_reset(model):
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_normal_(m.weight)
            nn.init.constant_(m.bias, 0)


class DownConv(nn.Module):
    def __init__(self, in_channels, out_channels, blocks=0, pool=0):
        super(DownConv, self).__init__()
        self.conv1 = MeshConv(in_channels, out_channels)
        self.conv2 = nn.ModuleList([MeshConv(out_channels, out_channels) for _ in range(blocks)])
        self.bn = nn.ModuleList([nn.InstanceNorm2d(out_channels) for _ in range(blocks + 1)])
        self.pool = MeshPool(pool) if pool else None

    def forward_em(self, x, batch):
        x1 = _norm_em(self.bn[0], self.conv1.forward_em(x, batch), post_relu=True)
        for idx, conv in enumerate(self.conv2):
            if RES_PASS:
                x2, res = conv.forward_em(x1, batch, pass_input=True)   # res is x1: both of its gradients meet in the conv's backward
            else:
                x2, res = conv.forward_em(x1, batch), x1
            x1 = _norm_em(self.bn[idx + 1], x2, post_add=res, post_relu=True)
        before_pool = None
        if self.pool is not None:
            before_pool = x1
            x1 = self.pool.forward_em(x1, batch)
        return x1, before_pool

    def forward(self, x):
        fe, meshes = x
        batch = MeshBatch.of(meshes, fe.shape[2], fe.device)
        out, before = self.forward_em(to_edge_major(fe), batch)
        return as_bce(out), (None if before is None else as_bce(before))


class UpConv(nn.Module):
    def __init__(self, in_channels, out_channels, blocks=0, unroll=0, residual=True, batch_norm=True, transfer_data=True):
        super(UpConv, self).__init__()
        self.residual, self.transfer_data = residual, transfer_data
        self.up_conv = MeshConv(in_channels, out_channels)
        self.conv1 = MeshConv(2 * out_channels if transfer_data else out_channels, out_channels)
        self.conv2 = nn.ModuleList([MeshConv(out_channels, out_channels) for _ in range(blocks)])
        self.bn = nn.ModuleList([nn.InstanceNorm2d(out_channels) for _ in range(blocks + 1)]) if batch_norm else []
        self.unroll = MeshUnpool(unroll) if unroll else None

    def forward_em(self, x, batch, from_down=None):
        x1 = self.up_conv.forward_em(x, batch)
        if self.unroll is not None:
            x1 = self.unroll.forward_em(x1, batch)
        if self.transfer_data:
            x1 = torch.cat((x1, from_down), 2)
        x1 = self.conv1.forward_em(x1, batch)
        x1 = _norm_em(self.bn[0], x1, post_relu=True) if len(self.bn) else F.relu(x1)
        for idx, conv in enumerate(self.conv2):
            if self.residual and len(self.bn) and RES_PASS:
                x2, res = conv.forward_em(x1, batch, pass_input=True)   # res is x1: both of its gradients meet in the conv's backward
            else:
                x2 = conv.forward_em(x1, batch)
                res = x1 if self.residual else None
            if len(self.bn):
                x1 = _norm_em(self.bn[idx + 1], x2, post_add=res, post_relu=True)
            else:
                x1 = F.relu(x2 if res is None else x2 + res)
        return x1

    def forward(self, x, from_down=None):
        from_up, meshes = x
        batch = MeshBatch.of(meshes, from_up.shape[2], from_up.device)
        fd = None if from_down is None else to_edge_major(from_down)
        return as_bce(self.forward_em(to_edge_major(from_up), batch, fd))


class MeshEncoder(nn.Module):
    """Down-convolution stack, optionally followed by the fully-connected head of networks.py:304-323 (`fcs` layer widths,
    `global_pool` 'max' / 'avg' over the last pooled resolution or none = flatten [C x E])."""

    def __init__(self, pools, convs, fcs=None, blocks=0, global_pool=None):
        super(MeshEncoder, self).__init__()
        self.fcs = None
        self.global_pool = None
        self.convs = nn.ModuleList([DownConv(convs[i], convs[i + 1], blocks=blocks,
                                             pool=pools[i + 1] if i + 1 < len(pools) else 0)
                                    for i in range(len(convs) - 1)])
        if fcs is not None:
            last_length = convs[-1]
            if global_pool is not None:
                if global_pool == 'max':
                    self.global_pool = nn.MaxPool1d(pools[-1])
                elif global_pool == 'avg':
                    self.global_pool = nn.AvgPool1d(pools[-1])
                else:
                    assert False, 'global_pool %s is not defined' % global_pool
            else:
                last_length *= pools[-1]
            fcs = list(fcs)
            if fcs[0] == last_length:
                fcs = fcs[1:]
            layers, norms = [], []
            for length in fcs:
                layers.append(nn.Linear(last_length, length))
                norms.append(nn.InstanceNorm1d(length))
                last_length = length
            self.fcs = nn.ModuleList(layers)
            self.fcs_bn = nn.ModuleList(norms)
        _xavier_reset(self)

    def _head(self, fe):
        """fe: logical [B, C, E] (a view of the edge-major tensor)."""
        if self.global_pool is not None:
            fe = self.global_pool(fe)
        fe = fe.contiguous().view(fe.size()[0], -1)
        for i in range(len(self.fcs)):
            fe = self.fcs[i](fe)
            if self.fcs_bn:
                fe = self.fcs_bn[i](fe.unsqueeze(1)).squeeze(1)
            if i < len(self.fcs) - 1:
                fe = F.relu(fe)
        return fe

    def forward_em(self, fe, batch):
        """-> (edge-major [B, E, C] features, or the [B, fcs[-1]] code when the head is present; encoder skips)."""
        outs = []
        for conv in self.convs:
            fe, before = conv.forward_em(fe, batch)
            outs.append(before)
        if self.fcs is not None:
            fe = self._head(as_bce(fe))
        return fe, outs

    def forward(self, x):
        fe, meshes = x
        batch = MeshBatch.of(meshes, fe.shape[2], fe.device)
        fe, outs = self.forward_em(to_edge_major(fe), batch)
        return (fe if self.fcs is not None else as_bce(fe)), [None if o is None else as_bce(o) for o in outs]

    def __call__(self, x):
        return self.forward(x)


class MeshDecoder(nn.Module):
    def __init__(self, unrolls, convs, blocks=0, batch_norm=True, transfer_data=True):
        super(MeshDecoder, self).__init__()
        self.up_convs = nn.ModuleList([UpConv(convs[i], convs[i + 1], blocks=blocks,
                                              unroll=unrolls[i] if i < len(unrolls) else 0,
                                              batch_norm=batch_norm, transfer_data=transfer_data)
                                       for i in range(len(convs) - 2)])
        self.final_conv = UpConv(convs[-2], convs[-1], blocks=blocks, unroll=False, batch_norm=batch_norm,
                                 transfer_data=False)
        _xavier_reset(self)

    def forward_em(self, fe, batch, encoder_outs=None):
        for i, up in enumerate(self.up_convs):
            fe = up.forward_em(fe, batch, None if encoder_outs is None else encoder_outs[-(i + 2)])
        return self.final_conv.forward_em(fe, batch)

    def forward(self, x, encoder_outs=None):
        fe, meshes = x
        batch = MeshBatch.of(meshes, fe.shape[2], fe.device)
        eo = None if encoder_outs is None else [None if o is None else to_edge_major(o) for o in encoder_outs]
        return as_bce(self.forward_em(to_edge_major(fe), batch, eo))


class MeshEncoderDecoder(nn.Module):
    """Network for fully-convolutional tasks (segmentation)."""

    def __init__(self, pools, down_convs, up_convs, blocks=0, transfer_data=True):
        super(MeshEncoderDecoder, self).__init__()
        self.transfer_data = transfer_data
        self.encoder = MeshEncoder(pools, down_convs, blocks=blocks)
        unrolls = list(pools[:-1])
        unrolls.reverse()
        self.decoder = MeshDecoder(unrolls, up_convs, blocks=blocks, transfer_data=transfer_data)

    def forward(self, x, meshes):
        batch = MeshBatch.of(meshes, x.shape[2], x.device)
        fe, before_pool = self.encoder.forward_em(to_edge_major(x), batch)
        fe = self.decoder.forward_em(fe, batch, before_pool)
        return as_bce(fe)
