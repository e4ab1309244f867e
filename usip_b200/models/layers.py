"""Host-side mirror of the hot-path layer types of models/layers.py.

These classes exist so that (a) `state_dict()` keys / shapes / init match the reference exactly
(`*.conv.{weight,bias}`, `*.norm.{weight,bias,running_mean,running_var,num_batches_tracked}`) and
(b) the reference's constructor signatures keep working.  They hold parameters; the arithmetic of a whole
network is scheduled by usip_b200/engine.py as a fused plan (point-major activations, BN folded into GEMM
prologues/epilogues), so the per-layer `forward` here is the stand-alone entry point that routes one layer
through the same kernels (usip_layer_fwd + usip_bn_finalize) for callers that use a layer on its own.

Reference: EquivariantLayer layers.py:248-303, MyConv2d :172-216, MyBatchNorm1d/2d :23-121,
PointNet :524-544, GeneralKNNFusionModule :375-440.
"""
import math

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from usip_b200 import ops

_SUPPORTED_ACT = (None, "relu")


class _MyBatchNorm(_BatchNorm):
    """BatchNorm parameter/buffer holder with the reference's epoch-scheduled momentum (layers.py:62-66)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, momentum_decay_step=None, momentum_decay=1):
        super().__init__(num_features, eps, momentum, affine)
        self.momentum_decay_step = momentum_decay_step
        self.momentum_decay = momentum_decay
        self.momentum_original = self.momentum

    def _check_input_dim(self, input):
        return

    def scheduled_momentum(self, epoch=None):
        if (epoch is not None) and (epoch >= 1) and (self.momentum_decay_step is not None) and (self.momentum_decay_step > 0):
            self.momentum = self.momentum_original * (self.momentum_decay ** (epoch // self.momentum_decay_step))
            if self.momentum < 0.01:
                self.momentum = 0.01
        return self.momentum


class MyBatchNorm1d(_MyBatchNorm):
    pass


class MyBatchNorm2d(_MyBatchNorm):
    pass


def _check_cfg(activation, normalization):
    if activation not in _SUPPORTED_ACT:
        raise NotImplementedError("activation %r: the B200 path implements 'relu' / None (every shipped config, "
                                  "kitti|oxford|modelnet options: activation='relu')" % (activation,))
    if normalization not in (None, "batch"):
        raise NotImplementedError("normalization %r: the B200 path implements 'batch' / None" % (normalization,))


def _standalone_layer_forward(layer, x, epoch):
    """One conv1x1(+BN)(+ReLU) through the fused kernels, reference (B,C,*spatial) layout in and out.
    Forward only (the autograd-aware path is the network-level plan)."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in layer.parameters()):
        raise NotImplementedError("stand-alone layer autograd is not provided; differentiate through the "
                                  "network-level modules (RPN_Detector / DescriptorLiteOld)")
    B, C = x.shape[0], x.shape[1]
    spatial = x.shape[2:]
    rows = x.reshape(B, C, -1).permute(0, 2, 1).reshape(-1, C).contiguous()
    P = rows.shape[0]
    W = layer.conv.weight.detach().reshape(layer.conv.weight.shape[0], -1)
    Cout = W.shape[0]
    dev = x.device
    Y = torch.empty((P, Cout), dtype=torch.float32, device=dev)
    has_bn = layer.normalization == "batch"
    nt = ops.stat_slots(P, Cout, 0)
    part = torch.empty((nt, 2, Cout), dtype=torch.float32, device=dev) if (has_bn and layer.training) else None
    ops.layer_fwd(rows, W, layer.conv.bias.detach(), P, C, Cout, Y=Y, stat_partial=part)
    relu = layer.activation == "relu"
    if has_bn or relu:
        scale = torch.ones(Cout, dtype=torch.float32, device=dev)
        shift = torch.zeros(Cout, dtype=torch.float32, device=dev)
        if has_bn:
            n = layer.norm
            if layer.training:
                ops.bn_finalize(part, nt, P, Cout, n.weight.detach(), n.bias.detach(), n.eps, n.scheduled_momentum(epoch),
                                n.running_mean, n.running_var, scale, shift)
            else:
                ops.bn_eval_affine(n.weight.detach(), n.bias.detach(), n.running_mean, n.running_var, n.eps, scale, shift)
        # apply the folded affine (+ReLU) with the same kernel: identity weight is wasteful, so use the
        # group-select epilogue with gmax == gmin == Y, which evaluates relu(scale*y+shift) when relu is on.
        out = torch.empty_like(Y)
        if relu:
            ops.group_select(Y, Y, scale, shift, out, P, Cout)
        else:
            out = Y * scale + shift    # affine-only layers do not occur on the hot path
        Y = out
    return Y.reshape(B, -1, Cout).permute(0, 2, 1).reshape((B, Cout) + tuple(spatial)).contiguous()


class EquivariantLayer(nn.Module):
    """Conv1d(k=1) + optional BatchNorm1d + optional ReLU (layers.py:248-303)."""

    def __init__(self, num_in_channels, num_out_channels, activation='relu', normalization=None, momentum=0.1,
                 bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        _check_cfg(activation, normalization)
        self.num_in_channels = num_in_channels
        self.num_out_channels = num_out_channels
        self.activation = activation
        self.normalization = normalization
        self.conv = nn.Conv1d(num_in_channels, num_out_channels, kernel_size=1, stride=1, padding=0)
        if normalization == 'batch':
            self.norm = MyBatchNorm1d(num_out_channels, momentum=momentum, affine=True,
                                      momentum_decay_step=bn_momentum_decay_step, momentum_decay=bn_momentum_decay)
        self.weight_init()

    def weight_init(self):                                   # layers.py:278-287
        n = self.conv.kernel_size[0] * self.conv.in_channels
        self.conv.weight.data.normal_(0, math.sqrt(2. / n))
        self.conv.bias.data.fill_(0)
        if self.normalization == 'batch':
            self.norm.weight.data.fill_(1)
            self.norm.bias.data.zero_()

    def forward(self, x, epoch=None):
        return _standalone_layer_forward(self, x, epoch)


class MyConv2d(nn.Module):
    """Conv2d(1x1) + optional BatchNorm2d + optional ReLU (layers.py:172-216)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, activation=None,
                 normalization=None, momentum=0.1, bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        _check_cfg(activation, normalization)
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if tuple(ks) != (1, 1) or stride != 1 or padding != 0 or not bias:
            raise NotImplementedError("only the hot path's 1x1 / stride 1 / pad 0 / bias conv is implemented")
        self.activation = activation
        self.normalization = normalization
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        if normalization == 'batch':
            self.norm = MyBatchNorm2d(out_channels, momentum=momentum, affine=True,
                                      momentum_decay_step=bn_momentum_decay_step, momentum_decay=bn_momentum_decay)
        self.weight_init()

    def weight_init(self):                                   # layers.py:196-205
        n = self.conv.kernel_size[0] * self.conv.kernel_size[1] * self.conv.in_channels
        self.conv.weight.data.normal_(0, math.sqrt(2. / n))
        self.conv.bias.data.fill_(0)
        if self.normalization == 'batch':
            self.norm.weight.data.fill_(1)
            self.norm.bias.data.zero_()

    def forward(self, x, epoch=None):
        return _standalone_layer_forward(self, x, epoch)


class PointNet(nn.Module):
    """Stack of EquivariantLayers; all but the last have BN+act (layers.py:524-544)."""

    def __init__(self, in_channels, out_channels_list, activation, normalization, momentum=0.1,
                 bn_momentum_decay_step=None, bn_momentum_decay=1, output_init_radius=None):
        super().__init__()
        self.layers = nn.ModuleList()
        prev = in_channels
        for i, c_out in enumerate(out_channels_list):
            if i != len(out_channels_list) - 1:
                self.layers.append(EquivariantLayer(prev, c_out, activation, normalization, momentum,
                                                    bn_momentum_decay_step, bn_momentum_decay))
            else:
                self.layers.append(EquivariantLayer(prev, c_out, None, None))
            prev = c_out
        if output_init_radius is not None:
            self.layers[len(out_channels_list) - 1].conv.bias.data.uniform_(-1 * output_init_radius, output_init_radius)

    def forward(self, x, epoch=None):
        for layer in self.layers:
            x = layer(x, epoch)
        return x


class GeneralKNNFusionModule(nn.Module):
    """Parameter container of the kNN fusion block (layers.py:375-440); scheduled by engine.detector_forward."""

    def __init__(self, in_channels, out_channels_list_before, out_channels_list_after, activation, normalization,
                 momentum=0.1, bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        if activation != 'relu' or normalization != 'batch':
            raise NotImplementedError("GeneralKNNFusionModule: the fused plan needs relu + batch norm on every layer "
                                      "(layers.py:383-399 with the shipped options)")
        self.layers_before = nn.ModuleList()
        prev = in_channels
        for c_out in out_channels_list_before:
            self.layers_before.append(MyConv2d(prev, c_out, kernel_size=1, stride=1, padding=0, bias=True,
                                               activation=activation, normalization=normalization, momentum=momentum,
                                               bn_momentum_decay_step=bn_momentum_decay_step,
                                               bn_momentum_decay=bn_momentum_decay))
            prev = c_out
        self.layers_after = nn.ModuleList()
        prev = 2 * prev
        for c_out in out_channels_list_after:
            self.layers_after.append(MyConv2d(prev, c_out, kernel_size=1, stride=1, padding=0, bias=True,
                                              activation=activation, normalization=normalization, momentum=momentum,
                                              bn_momentum_decay_step=bn_momentum_decay_step,
                                              bn_momentum_decay=bn_momentum_decay))
            prev = c_out

    def forward(self, query, database, x, K, epoch=None):
        raise NotImplementedError("GeneralKNNFusionModule is executed inside RPN_Detector's fused plan "
                                  "(usip_b200.engine.detector_forward); it has no stand-alone forward")
