"""The BASELINE.json model workloads, built from ``brevitas_b200.nn`` layers with the reference's quantizers.

* ``tfc``          bnn_pynq TFC (src/brevitas_examples/bnn_pynq/models/FC.py:19-69, common.py:16-42, cfg/tfc_2w2a.ini)
* ``resnet18``     torchvision-topology ResNet-18 with every Conv2d -> QuantConv2d (default Int8WeightPerTensorFloat,
                   nn/quant_conv.py:129), every ReLU -> QuantReLU (default Uint8ActPerTensorFloat,
                   nn/quant_activation.py:18), Linear -> QuantLinear (the reference has no ResNet-18)
* ``mobilenet_v1`` src/brevitas_examples/imagenet_classification/models/mobilenetv1.py:76-184 with
                   CommonIntWeightPerChannelQuant / CommonUintActQuant (models/common.py:10-47), QuantTensor-carrying
                   activations, the truncating ``QuantAvgPool2d`` and the ``IntBias`` classifier (accumulator scale and
                   bit-width from the input QuantTensor), as in the reference.
"""
from functools import reduce
from operator import mul

import torch
from torch import nn

from brevitas_b200.fused_bn import bn_act_quant
from brevitas_b200.nn import QuantAvgPool2d, QuantConv2d, QuantIdentity, QuantLinear, QuantReLU
from brevitas_b200.quant import (ActQuantizer, IntBias, Int8WeightPerTensorFloat, Uint8ActPerTensorFloatMaxInit,
                                 WeightQuantizer)


# ---- bnn_pynq -------------------------------------------------------------------------------------------------
class _CommonQuant:
    """bnn_pynq/models/common.py:16-33: CONST scaling, FP restriction, narrow signed; 1 bit -> binary"""
    scaling_impl_type = "CONST"
    restrict_scaling_type = "FP"
    float_to_int_impl_type = "ROUND"
    scaling_per_output_channel = False
    narrow_range = True
    signed = True

    @classmethod
    def _quant_type(cls):
        if cls.bit_width is None:
            return "FP"
        return "BINARY" if cls.bit_width == 1 else "INT"


class CommonWeightQuant(_CommonQuant, WeightQuantizer):
    scaling_const = 1.0


class CommonActQuant(_CommonQuant, ActQuantizer):
    min_val = -1.0
    max_val = 1.0


class TensorNorm(nn.Module):
    """bnn_pynq/models/tensor_norm.py: batch-norm over the whole tensor (scalar statistics)"""

    def __init__(self, eps=1e-4, momentum=0.1):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.weight = nn.Parameter(torch.ones(1))
        self.bias = nn.Parameter(torch.zeros(1))
        self.register_buffer('running_mean', torch.zeros(1))
        self.register_buffer('running_var', torch.ones(1))

    def forward(self, x):
        if self.training:
            mean = x.mean()
            unbias_var = x.var(unbiased=True)
            biased_var = x.var(unbiased=False)
            # in place: a CUDA-graph replay must mutate the REGISTERED buffers (rebinding the attribute would leave
            # replays writing a pool tensor nobody reads; same values as tensor_norm.py's out-of-place form)
            with torch.no_grad():
                self.running_mean.mul_(1 - self.momentum).add_(self.momentum * mean.detach())
                self.running_var.mul_(1 - self.momentum).add_(self.momentum * unbias_var.detach())
            inv_std = 1 / (biased_var + self.eps).pow(0.5)
            return (x - mean) * inv_std * self.weight + self.bias
        return ((x - self.running_mean) / (self.running_var + self.eps).pow(0.5)) * self.weight + self.bias


class FC(nn.Module):
    """bnn_pynq FC (FC.py:19-69)"""
    DROPOUT = 0.2

    def __init__(self, num_classes=10, weight_bit_width=2, act_bit_width=2, in_bit_width=2, out_features=(64, 64, 64),
                 in_features=(28, 28)):
        super().__init__()
        self.features = nn.ModuleList()
        self.features.append(QuantIdentity(act_quant=CommonActQuant, bit_width=in_bit_width))
        self.features.append(nn.Dropout(p=self.DROPOUT))
        fin = reduce(mul, in_features)
        for fout in out_features:
            self.features.append(QuantLinear(fin, fout, bias=False, weight_bit_width=weight_bit_width,
                                             weight_quant=CommonWeightQuant))
            fin = fout
            self.features.append(nn.BatchNorm1d(num_features=fin))
            self.features.append(QuantIdentity(act_quant=CommonActQuant, bit_width=act_bit_width))
            self.features.append(nn.Dropout(p=self.DROPOUT))
        self.features.append(QuantLinear(fin, num_classes, bias=False, weight_bit_width=weight_bit_width,
                                         weight_quant=CommonWeightQuant))
        self.features.append(TensorNorm())
        for m in self.modules():
            if isinstance(m, QuantLinear):
                torch.nn.init.uniform_(m.weight.data, -1, 1)

    def clip_weights(self, min_val, max_val):
        for mod in self.features:
            if isinstance(mod, QuantLinear):
                mod.weight.data.clamp_(min_val, max_val)

    def forward(self, x):
        x = x.view(x.shape[0], -1)
        x = 2.0 * x - 1.0
        for mod in self.features:
            x = mod(x)
        return x


class SqrHingeLoss(nn.Module):
    """bnn_pynq/models/losses.py:9-31: mean(max(0, 1 - y*t)^2) with targets in {-1, +1}"""

    def forward(self, predictions, targets):
        out = (1. - predictions * targets).clamp_min(0.)
        return (out * out).mean()


def tfc(weight_bit_width=2, act_bit_width=2, in_bit_width=2):
    return FC(10, weight_bit_width, act_bit_width, in_bit_width)


# ---- ResNet-18 ------------------------------------------------------------------------------------------------
class _MirrorLayers:
    """layer namespace: this repository's mirror of brevitas.nn"""
    QuantConv2d, QuantReLU, QuantLinear = QuantConv2d, QuantReLU, QuantLinear


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, act_kw=None, L=_MirrorLayers, fuse_bn=False):
        super().__init__()
        act_kw = act_kw or {}
        self.fuse_bn = fuse_bn
        self.conv1 = L.QuantConv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = L.QuantReLU(**act_kw)
        self.conv2 = L.QuantConv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu2 = L.QuantReLU(**act_kw)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        if self.fuse_bn:          # batch-norm + ReLU + quantizer in fused passes (brevitas_b200/fused_bn.py)
            out = bn_act_quant(self.bn1, self.relu1, self.conv1(x))
        else:
            out = self.relu1(self.bn1(self.conv1(x)))
        if self.downsample is not None:
            identity = self.downsample(x)
        if self.fuse_bn:          # relu(bn(.) + identity) with its quantizer: the residual variant of the fused passes
            return bn_act_quant(self.bn2, self.relu2, self.conv2(out), residual=identity)
        out = self.bn2(self.conv2(out))
        return self.relu2(out + identity)


class ResNet18(nn.Module):
    """``L``: the layer namespace -- the mirror (default) or the reference's own ``brevitas.nn`` (qat/ref_models.py)"""

    def __init__(self, num_classes=1000, act_kw=None, L=_MirrorLayers, fuse_bn=False):
        super().__init__()
        act_kw = act_kw or {}
        self.L = L
        self.fuse_bn = fuse_bn
        self.conv1 = L.QuantConv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = L.QuantReLU(**act_kw)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.inplanes = 64
        self.layer1 = self._make_layer(64, 2, 1, act_kw)
        self.layer2 = self._make_layer(128, 2, 2, act_kw)
        self.layer3 = self._make_layer(256, 2, 2, act_kw)
        self.layer4 = self._make_layer(512, 2, 2, act_kw)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = L.QuantLinear(512, num_classes, True)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def _make_layer(self, planes, blocks, stride, act_kw):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(self.L.QuantConv2d(self.inplanes, planes, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample, act_kw, self.L, self.fuse_bn)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(planes, planes, act_kw=act_kw, L=self.L, fuse_bn=self.fuse_bn))
        return nn.Sequential(*layers)

    def forward(self, x):
        if self.fuse_bn:
            x = self.maxpool(bn_act_quant(self.bn1, self.relu, self.conv1(x)))
        else:
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(num_classes=1000, collect_stats_steps=300, fuse_bn=False):
    return ResNet18(num_classes, act_kw={"collect_stats_steps": collect_stats_steps}, fuse_bn=fuse_bn)


# ---- MobileNetV1 (imagenet_classification) --------------------------------------------------------------------
class CommonIntWeightPerTensorQuant(Int8WeightPerTensorFloat):
    """imagenet_classification/models/common.py:10-17"""
    scaling_min_val = 2e-16
    bit_width = None


class CommonIntWeightPerChannelQuant(CommonIntWeightPerTensorQuant):
    """models/common.py:19-24"""
    scaling_per_output_channel = True


class CommonUintActQuant(Uint8ActPerTensorFloatMaxInit):
    """models/common.py:39-47: learned LOG_FP scale initialised at 6.0"""
    scaling_min_val = 2e-16
    bit_width = None
    max_val = 6.0
    restrict_scaling_type = "LOG_FP"


FIRST_LAYER_BIT_WIDTH = 8


class ConvBlock(nn.Module):
    """mobilenetv1.py:76-115"""

    def __init__(self, in_channels, out_channels, kernel_size, weight_bit_width, act_bit_width, stride=1, padding=0,
                 groups=1, bn_eps=1e-5, activation_scaling_per_channel=False, fuse_bn=False):
        super().__init__()
        self.fuse_bn = fuse_bn
        self.conv = QuantConv2d(in_channels, out_channels, kernel_size, stride, padding, groups=groups, bias=False,
                                weight_quant=CommonIntWeightPerChannelQuant, weight_bit_width=weight_bit_width)
        self.bn = nn.BatchNorm2d(out_channels, eps=bn_eps)
        self.activation = QuantReLU(act_quant=CommonUintActQuant, bit_width=act_bit_width,
                                    per_channel_broadcastable_shape=(1, out_channels, 1, 1),
                                    scaling_per_output_channel=activation_scaling_per_channel,
                                    return_quant_tensor=True)

    def forward(self, x):
        if self.fuse_bn:
            return bn_act_quant(self.bn, self.activation, self.conv(x))
        return self.activation(self.bn(self.conv(x)))


class DwsConvBlock(nn.Module):
    """mobilenetv1.py:45-73"""

    def __init__(self, in_channels, out_channels, stride, bit_width, pw_activation_scaling_per_channel=False,
                 fuse_bn=False):
        super().__init__()
        self.dw_conv = ConvBlock(in_channels, in_channels, 3, bit_width, bit_width, stride, 1, groups=in_channels,
                                 fuse_bn=fuse_bn)
        self.pw_conv = ConvBlock(in_channels, out_channels, 1, bit_width, bit_width,
                                 activation_scaling_per_channel=pw_activation_scaling_per_channel, fuse_bn=fuse_bn)

    def forward(self, x):
        return self.pw_conv(self.dw_conv(x))


class MobileNetV1(nn.Module):
    """mobilenetv1.py:118-166 (4-bit by default, first layer 8-bit)"""

    def __init__(self, bit_width=4, num_classes=1000, width_scale=1.0, fuse_bn=False):
        super().__init__()
        channels = [[32], [64], [128, 128], [256, 256], [512] * 6, [1024, 1024]]
        if width_scale != 1.0:
            channels = [[int(c * width_scale) for c in ci] for ci in channels]
        self.features = nn.Sequential()
        cin = channels[0][0]
        self.features.add_module('init_block', ConvBlock(3, cin, 3, FIRST_LAYER_BIT_WIDTH, bit_width, stride=2,
                                                         activation_scaling_per_channel=True, fuse_bn=fuse_bn))
        for i, stage_channels in enumerate(channels[1:]):
            stage = nn.Sequential()
            per_channel = i < len(channels[1:]) - 1
            for j, cout in enumerate(stage_channels):
                stride = 2 if (j == 0 and i != 0) else 1
                stage.add_module(f'unit{j + 1}', DwsConvBlock(cin, cout, stride, bit_width, per_channel, fuse_bn))
                cin = cout
            self.features.add_module(f'stage{i + 1}', stage)
        self.final_pool = QuantAvgPool2d(kernel_size=7, stride=1, bit_width=bit_width)
        self.output = QuantLinear(cin, num_classes, bias=True, bias_quant=IntBias,
                                  weight_quant=CommonIntWeightPerTensorQuant, weight_bit_width=bit_width)

    def forward(self, x):
        x = self.final_pool(self.features(x))
        return self.output(x.view(x.size(0), -1))


def mobilenet_v1(bit_width=4, num_classes=1000, fuse_bn=False):
    return MobileNetV1(bit_width, num_classes, fuse_bn=fuse_bn)
