"""CPU oracle networks -- TEST INFRASTRUCTURE (see oracle/meshcnn_oracle.py for the rules).

Restates models/networks.py:124-388 of the reference on top of the oracle layers, with the same module tree so a
state dict moves freely between the reference, this oracle and meshcnn_b200.networks.  Pinned to the reference by
tests/golden/net_*.npz.  ``dense=True`` makes MeshPool use the reference's dense n x n group matrix and MeshConv the
reference's op sequence, which is what bench.py's CPU baseline times.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import meshcnn_oracle as O


class OMeshConv(nn.Module):                                   # mesh_conv.py:5-26
    def __init__(self, cin, cout, bias=True, cost_faithful=False):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=(1, 5), bias=bias)
        self.cost_faithful = cost_faithful

    def forward(self, x, meshes):
        fn = O.conv_forward_reference_cost if self.cost_faithful else O.conv_forward
        return fn(x, meshes, self.conv.weight, self.conv.bias)


class OMeshPool(nn.Module):                                   # mesh_pool.py:9-39
    def __init__(self, target, dense=False):
        super().__init__()
        self.target, self.dense = target, dense
        self.norm_override = None
        self.last_logs = None

    def forward(self, x, meshes):
        out, self.last_logs = O.pool_forward(x, meshes, self.target, norms=self.norm_override, dense=self.dense)
        return out


class OMeshUnpool(nn.Module):                                 # mesh_unpool.py:6-41
    def __init__(self, target):
        super().__init__()
        self.target = target

    def forward(self, x, meshes):
        return O.unpool_forward(x, meshes, self.target)


class OMResConv(nn.Module):                                   # networks.py:159-179
    def __init__(self, cin, cout, skips, cf):
        super().__init__()
        self.skips = skips
        self.conv0 = OMeshConv(cin, cout, bias=False, cost_faithful=cf)
        for i in range(skips):
            setattr(self, 'bn%d' % (i + 1), nn.BatchNorm2d(cout))
            setattr(self, 'conv%d' % (i + 1), OMeshConv(cout, cout, bias=False, cost_faithful=cf))

    def forward(self, x, meshes):
        x = self.conv0(x, meshes)
        x1 = x
        for i in range(self.skips):
            x = getattr(self, 'bn%d' % (i + 1))(F.relu(x))
            x = getattr(self, 'conv%d' % (i + 1))(x, meshes)
        return F.relu(x + x1)


class OMeshConvNet(nn.Module):                                # networks.py:124-157
    def __init__(self, norm_layer, nf0, conv_res, nclasses, input_res, pool_res, fc_n, nresblocks=3, dense=False):
        super().__init__()
        self.k = [nf0] + conv_res
        self.res = [input_res] + pool_res
        for i in range(len(self.k) - 1):
            setattr(self, 'conv%d' % i, OMResConv(self.k[i], self.k[i + 1], nresblocks, dense))
            setattr(self, 'norm%d' % i, norm_layer(self.k[i + 1]))
            setattr(self, 'pool%d' % i, OMeshPool(self.res[i + 1], dense))
        self.gp = nn.AvgPool1d(self.res[-1])
        self.fc1 = nn.Linear(self.k[-1], fc_n)
        self.fc2 = nn.Linear(fc_n, nclasses)

    def forward(self, x, meshes):
        for i in range(len(self.k) - 1):
            x = getattr(self, 'conv%d' % i)(x, meshes)
            x = F.relu(getattr(self, 'norm%d' % i)(x))
            x = getattr(self, 'pool%d' % i)(x, meshes)
        x = self.gp(x).view(-1, self.k[-1])
        return self.fc2(F.relu(self.fc1(x)))

    def pools(self):
        return [getattr(self, 'pool%d' % i) for i in range(len(self.k) - 1)]


class ODownConv(nn.Module):                                   # networks.py:201-239
    def __init__(self, cin, cout, blocks, pool, dense):
        super().__init__()
        self.conv1 = OMeshConv(cin, cout, cost_faithful=dense)
        self.conv2 = nn.ModuleList([OMeshConv(cout, cout, cost_faithful=dense) for _ in range(blocks)])
        self.bn = nn.ModuleList([nn.InstanceNorm2d(cout) for _ in range(blocks + 1)])
        self.pool = OMeshPool(pool, dense) if pool else None

    def forward(self, fe, meshes):
        x1 = F.relu(self.bn[0](self.conv1(fe, meshes)))
        for i, conv in enumerate(self.conv2):
            x1 = F.relu(self.bn[i + 1](conv(x1, meshes)) + x1)
        x1 = x1.squeeze(3)
        before = None
        if self.pool is not None:
            before = x1
            x1 = self.pool(x1, meshes)
        return x1, before


class OUpConv(nn.Module):                                     # networks.py:242-290
    def __init__(self, cin, cout, blocks, unroll, transfer, dense):
        super().__init__()
        self.transfer = transfer
        self.up_conv = OMeshConv(cin, cout, cost_faithful=dense)
        self.conv1 = OMeshConv(2 * cout if transfer else cout, cout, cost_faithful=dense)
        self.conv2 = nn.ModuleList([OMeshConv(cout, cout, cost_faithful=dense) for _ in range(blocks)])
        self.bn = nn.ModuleList([nn.InstanceNorm2d(cout) for _ in range(blocks + 1)])
        self.unroll = OMeshUnpool(unroll) if unroll else None

    def forward(self, fe, meshes, from_down=None):
        x1 = self.up_conv(fe, meshes).squeeze(3)
        if self.unroll is not None:
            x1 = self.unroll(x1, meshes)
        if self.transfer:
            x1 = torch.cat((x1, from_down), 1)
        x1 = F.relu(self.bn[0](self.conv1(x1, meshes)))
        for i, conv in enumerate(self.conv2):
            x1 = F.relu(self.bn[i + 1](conv(x1, meshes)) + x1)
        return x1.squeeze(3)


class _Enc(nn.Module):
    def __init__(self, pools, convs, blocks, dense):
        super().__init__()
        self.convs = nn.ModuleList([ODownConv(convs[i], convs[i + 1], blocks, pools[i + 1] if i + 1 < len(pools) else 0, dense)
                                    for i in range(len(convs) - 1)])


class _Dec(nn.Module):
    def __init__(self, unrolls, convs, blocks, dense):
        super().__init__()
        self.up_convs = nn.ModuleList([OUpConv(convs[i], convs[i + 1], blocks, unrolls[i] if i < len(unrolls) else 0, True, dense)
                                       for i in range(len(convs) - 2)])
        self.final_conv = OUpConv(convs[-2], convs[-1], blocks, 0, False, dense)


class OMeshEncoderDecoder(nn.Module):                         # networks.py:182-199, :293-376
    def __init__(self, pools, down_convs, up_convs, blocks=0, dense=False):
        super().__init__()
        self.encoder = _Enc(pools, down_convs, blocks, dense)
        unrolls = list(pools[:-1])[::-1]
        self.decoder = _Dec(unrolls, up_convs, blocks, dense)

    def forward(self, x, meshes):
        outs = []
        fe = x
        for conv in self.encoder.convs:
            fe, before = conv(fe, meshes)
            outs.append(before)
        for i, up in enumerate(self.decoder.up_convs):
            fe = up(fe, meshes, outs[-(i + 2)])
        return self.decoder.final_conv(fe, meshes)

    def pools(self):
        return [c.pool for c in self.encoder.convs if c.pool is not None]
