"""``DCNv2P`` / ``DCNv2P2`` — the conv layers the BEVFormer backbones (ResNet-101-DCN stages 3-4) instantiate in the
plugin configs, mirroring det2trt/models/modules/cnn/dcn.py:31-164: a modulated deformable convolution that "acts as a
normal conv layer" — its own ``conv_offset`` (a plain conv producing ``deform_groups * 3 * kh * kw`` channels), then
``offset = cat(o1, o2)``, ``mask = sigmoid(mask)`` (:70-74) and the plugin op (:75-86).

The reference derives from mmcv's ``ModulatedDeformConv2d`` (absent here); this class restates that base in plain
``torch.nn`` with the same constructor arguments, parameter names (``weight``, ``bias``, ``conv_offset.weight``,
``conv_offset.bias``: checkpoints load unchanged, including the pre-version-2 key names, :88-128) and initialisation
(kaiming-uniform-like ``weight``, zero ``bias``, zero ``conv_offset``, :62-66). Layer names are registered in a small
name -> class table like mmcv's CONV_LAYERS registry (:31, :131)."""
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from ...functions import modulated_deformable_conv2d, modulated_deformable_conv2d2

CONV_LAYERS = {}


def _register(name):
    def deco(cls):
        CONV_LAYERS[name] = cls
        return cls

    return deco


@_register("DCNv2P")
class ModulatedDeformConv2dPackPlugin(nn.Module):
    """Args as ``nn.Conv2d``: in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
    deform_groups=1, bias=True (mmcv ModulatedDeformConv2d's signature; ``bias='auto'`` is resolved by the caller)."""

    _version = 2
    _op = staticmethod(modulated_deformable_conv2d)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deform_groups=1,
                 bias=True):  # fmt: skip
        super().__init__()
        if in_channels % groups or out_channels % groups or in_channels % deform_groups:
            raise ValueError("in_channels / out_channels must be divisible by groups, in_channels by deform_groups")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = (_pair(kernel_size), _pair(stride), _pair(padding),
                                                                      _pair(dilation))  # fmt: skip
        self.groups, self.deform_groups = groups, deform_groups
        self.transposed, self.output_padding = False, _pair(0)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=self.stride, padding=self.padding,
                                     dilation=self.dilation, bias=True)  # fmt: skip
        self.init_weights()

    def init_weights(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()
        if hasattr(self, "conv_offset"):
            self.conv_offset.weight.data.zero_()
            self.conv_offset.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return self._op(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups,
                        self.deform_groups)  # fmt: skip

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        version = local_metadata.get("version", None)
        # early checkpoints: "<name>_offset.*" instead of "<name>.conv_offset.*" (dcn.py:88-128). Recent torch hands this
        # method a prefix-filtered view of the state dict, so the old names are only visible when loading this layer
        # directly (or through mmcv-style loaders that pass the full dict), exactly as for the reference class.
        if version is None or version < 2:
            for leaf in ("weight", "bias"):
                old, new = prefix[:-1] + "_offset." + leaf, prefix + "conv_offset." + leaf
                if new not in state_dict and old in state_dict:
                    state_dict[new] = state_dict.pop(old)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


@_register("DCNv2P2")
class ModulatedDeformConv2dPackPlugin2(ModulatedDeformConv2dPackPlugin):
    """Same layer bound to ``modulated_deformable_conv2d2`` (plugin ModulatedDeformableConv2dTRT2, dcn.py:131-164)."""

    _op = staticmethod(modulated_deformable_conv2d2)
