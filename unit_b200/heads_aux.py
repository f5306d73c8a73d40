"""Box heads and mask heads registered under the reference's names.

``Res5BoxHead`` / ``Res5BoxHeadWithMask`` (modeling/roi_heads/box_head.py:47-89,137-141) are dense cuDNN
convolutions and OUT OF SCOPE for the hand-written path: they are rebuilt here from stock ``torch.nn`` modules with
Detectron2's parameter names (``res5.{0,1,2}.{conv1,conv2,conv3,shortcut}.{weight,norm.*}``) so that checkpoints
load and the YAMLs build.  The mask heads (modeling/roi_heads/mask_head.py) keep their deconv / 1x1 convs in stock
PyTorch and run the base->novel mask transfer + class select + sigmoid as one fused kernel.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .predictors import _freeze
from .registry import ROI_BOX_HEAD_REGISTRY, ROI_MASK_HEAD_REGISTRY, configurable
from .structures import Instances, ShapeSpec


class FrozenBatchNorm2d(nn.Module):
    """[D2] FrozenBatchNorm2d: fixed statistics and affine (RESNETS.NORM = "FrozenBN")."""

    def __init__(self, num_features: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        bias = self.bias - self.running_mean * scale
        return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + bias.reshape(1, -1, 1, 1).to(x.dtype)


class _ConvNorm(nn.Conv2d):
    def __init__(self, cin, cout, k, stride=1, padding=0, groups=1):
        super().__init__(cin, cout, k, stride=stride, padding=padding, groups=groups, bias=False)
        self.norm = FrozenBatchNorm2d(cout)

    def forward(self, x):
        return self.norm(super().forward(x))


class BottleneckBlock(nn.Module):
    """[D2] ResNet BottleneckBlock (1x1, 3x3, 1x1 + projection shortcut)."""

    def __init__(self, cin, cout, *, bottleneck_channels, stride=1, num_groups=1, stride_in_1x1=False):
        super().__init__()
        self.shortcut = _ConvNorm(cin, cout, 1, stride=stride) if cin != cout else None
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = _ConvNorm(cin, bottleneck_channels, 1, stride=s1)
        self.conv2 = _ConvNorm(bottleneck_channels, bottleneck_channels, 3, stride=s3, padding=1, groups=num_groups)
        self.conv3 = _ConvNorm(bottleneck_channels, cout, 1)

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return F.relu_(out + sc)


@ROI_BOX_HEAD_REGISTRY.register()
class Res5BoxHead(nn.Module):
    """box_head.py:47-89: res5 stage (3 bottlenecks, first with stride 2) + global mean -> [R, 2048]."""

    def __init__(self, cfg, input_shape):
        super().__init__()
        r = cfg.MODEL.RESNETS
        factor = 2 ** 3
        bottleneck = r.NUM_GROUPS * r.WIDTH_PER_GROUP * factor
        out_channels = r.RES2_OUT_CHANNELS * factor
        blocks, cin = [], out_channels // 2
        for i, stride in enumerate([2, 1, 1]):
            blocks.append(BottleneckBlock(cin, out_channels, bottleneck_channels=bottleneck, stride=stride,
                                          num_groups=r.NUM_GROUPS, stride_in_1x1=r.STRIDE_IN_1X1))
            cin = out_channels
        self.res5 = nn.Sequential(*blocks)
        self.out_channels = out_channels

    def forward(self, x):
        return self.res5(x).mean(dim=[2, 3])

    @property
    def output_shape(self):
        return ShapeSpec(channels=self.out_channels, height=1, width=1)


@ROI_BOX_HEAD_REGISTRY.register()
class Res5BoxHeadWithMask(Res5BoxHead):
    """box_head.py:137-141: keeps the 7x7 map (the heads mean-pool it for the box branch); ``output_shape`` is
    inherited, i.e. still (2048, 1, 1), which is what sizes the predictor and the mask head in the reference."""

    def forward(self, x):
        return self.res5(x)


class _MaskHeadBase(nn.Module):
    """[D2] MaskRCNNConvUpsampleHead layout: [3x3 conv + relu] * NUM_CONV, 2x2 stride-2 deconv + relu, 1x1 predictor."""

    @configurable
    def __init__(self, input_shape, *, num_classes, conv_dims, conv_norm="", vis_period=0, freeze_layers=()):
        super().__init__()
        assert len(conv_dims) >= 1 and not conv_norm
        self.vis_period = vis_period
        self.num_classes = num_classes
        self.conv_norm_relus = []
        cur = input_shape.channels
        for k, dim in enumerate(conv_dims[:-1]):
            conv = nn.Conv2d(cur, dim, 3, stride=1, padding=1)
            self.add_module(f"mask_fcn{k + 1}", conv)
            self.conv_norm_relus.append(conv)
            cur = dim
        self.deconv = nn.ConvTranspose2d(cur, conv_dims[-1], kernel_size=2, stride=2, padding=0)
        self.deconv_relu = nn.ReLU()
        self.predictor = nn.Conv2d(conv_dims[-1], num_classes, kernel_size=1)
        for layer in self.conv_norm_relus + [self.deconv]:
            nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(layer.bias, 0)
        nn.init.normal_(self.predictor.weight, std=0.001)
        nn.init.constant_(self.predictor.bias, 0)
        self._extra_init(conv_dims[-1], num_classes)
        _freeze(self, freeze_layers)

    def _extra_init(self, dim, num_classes):
        pass

    @classmethod
    def from_config(cls, cfg, input_shape):
        mh = cfg.MODEL.ROI_MASK_HEAD
        return {
            "input_shape": input_shape,
            "conv_dims": [mh.CONV_DIM] * (mh.NUM_CONV + 1),
            "conv_norm": mh.NORM,
            "num_classes": 1 if mh.CLS_AGNOSTIC_MASK else cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "vis_period": cfg.VIS_PERIOD,
            "freeze_layers": cfg.MODEL.FREEZE_LAYERS.MASK_HEAD,
        }

    def _trunk(self, x):
        for conv in self.conv_norm_relus:
            x = F.relu(conv(x))
        return self.deconv_relu(self.deconv(x))

    def _logits(self, x):
        return self.predictor(self._trunk(x)), None

    def forward(self, x, instances: List[Instances], similarity=None, base_classes=None, novel_classes=None,
                spec: Optional[ops.TransferSpec] = None):
        """mask_head.py:16-37 / 72-94.  Inference only: transfer + select + sigmoid fused, ``pred_masks`` attached."""
        if self.training:
            raise NotImplementedError("mask_rcnn_loss is outside the scoped RoI stage (SURVEY.md section 8a row a11)")
        fixed, delta = self._logits(x)
        s_seg = None if similarity is None else similarity["seg"]
        if spec is None:
            spec = ops.TransferSpec(self.num_classes, base_classes.tolist(), novel_classes.tolist(), x.device)
        cls = torch.cat([i.pred_classes for i in instances])
        _, probs = ops.mask_transfer(fixed, s_seg, spec, delta, cls, want_logits=False, want_probs=True)
        for prob, inst in zip(probs.split([len(i) for i in instances], dim=0), instances):
            inst.pred_masks = prob
        return instances


@ROI_MASK_HEAD_REGISTRY.register()
class MaskRCNNConvUpsampleHeadWithSimilarity(_MaskHeadBase):
    """mask_head.py:14-37."""


@ROI_MASK_HEAD_REGISTRY.register()
class MaskRCNNConvUpsampleHeadWithFineTune(_MaskHeadBase):
    """mask_head.py:39-94: adds the zero-initialised ``predictor_delta`` 1x1 conv."""

    def _extra_init(self, dim, num_classes):
        self.predictor_delta = nn.Conv2d(dim, num_classes, kernel_size=1)
        nn.init.constant_(self.predictor_delta.weight, 0.0)
        nn.init.constant_(self.predictor_delta.bias, 0.0)

    def _logits(self, x):
        t = self._trunk(x)
        return self.predictor(t), self.predictor_delta(t)

    def layers(self, x):
        return self._logits(x)
