"""Stand-in for the two third-party pieces mymodels.py imports from segmentation_models_pytorch:
`get_encoder("efficientnet-b4", ...)` and `decoders.unet.UnetDecoder` (mymodels.py:11-12, 242-258).

smp (and efficientnet_pytorch underneath) is an unpinned pip dependency of the reference
(requirements.txt:15) that is absent from /root/reference and not installable here, and the
authors ran a locally patched copy (`UnetDecoder_noise`, mymodels.py:12).  This module restates
the published smp 0.3.x / efficientnet_pytorch 0.7.x architecture from its documented layout —
module names, state_dict keys, channel plan (8,48,32,56,160,448) -> (256,128,64,32,16) — so the
head's wiring and parameter surface match; numerical identity with the authors' copy is
UNPINNED (no checkpoint or test vector of it exists in the reference).  It is <1 % of the path's
FLOPs and is left to stock PyTorch/cuDNN kernels: it is third-party code on the reference side
too, not part of the hand-written hot path.
"""
from __future__ import annotations

import math
import warnings
from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch import nn

# efficientnet_pytorch block strings for the B0 base, scaled by (width, depth)
_BASE_BLOCKS = [  # repeats, kernel, stride, expand, in, out, se
    (1, 3, 1, 1, 32, 16, 0.25), (2, 3, 2, 6, 16, 24, 0.25), (2, 5, 2, 6, 24, 40, 0.25),
    (3, 3, 2, 6, 40, 80, 0.25), (3, 5, 1, 6, 80, 112, 0.25), (4, 5, 2, 6, 112, 192, 0.25),
    (1, 3, 1, 6, 192, 320, 0.25)]
_PARAMS = {  # name: (width, depth, resolution, dropout)
    "efficientnet-b0": (1.0, 1.0, 224, 0.2), "efficientnet-b1": (1.0, 1.1, 240, 0.2),
    "efficientnet-b2": (1.1, 1.2, 260, 0.3), "efficientnet-b3": (1.2, 1.4, 300, 0.3),
    "efficientnet-b4": (1.4, 1.8, 380, 0.4), "efficientnet-b5": (1.6, 2.2, 456, 0.4),
}
_STAGE_IDXS = {"efficientnet-b0": (3, 5, 9, 16), "efficientnet-b1": (5, 8, 16, 23),
               "efficientnet-b2": (5, 8, 16, 23), "efficientnet-b3": (5, 8, 18, 26),
               "efficientnet-b4": (6, 10, 22, 32), "efficientnet-b5": (8, 13, 27, 39)}
_OUT_CHANNELS = {"efficientnet-b0": (3, 32, 24, 40, 112, 320), "efficientnet-b1": (3, 32, 24, 40, 112, 320),
                 "efficientnet-b2": (3, 32, 24, 48, 120, 352), "efficientnet-b3": (3, 40, 32, 48, 136, 384),
                 "efficientnet-b4": (3, 48, 32, 56, 160, 448), "efficientnet-b5": (3, 48, 40, 64, 176, 512)}


def _round_filters(f, width, divisor=8):
    f *= width
    new_f = max(divisor, int(f + divisor / 2) // divisor * divisor)
    if new_f < 0.9 * f:
        new_f += divisor
    return int(new_f)


def _round_repeats(r, depth):
    return int(math.ceil(depth * r))


class Conv2dStaticSamePadding(nn.Conv2d):
    """TF-style 'same' padding fixed at construction for a given image size (efficientnet_pytorch)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kwargs)
        ih, iw = (image_size, image_size) if isinstance(image_size, int) else image_size
        kh, kw = self.weight.shape[-2:]
        sh, sw = self.stride
        oh, ow = math.ceil(ih / sh), math.ceil(iw / sw)
        pad_h = max((oh - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        pad_w = max((ow - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0)
        if pad_h > 0 or pad_w > 0:
            self.static_padding = nn.ZeroPad2d((pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2))
        else:
            self.static_padding = nn.Identity()

    def forward(self, x):
        x = self.static_padding(x)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


def _out_size(size, stride):
    return int(math.ceil(size / stride))


class MBConvBlock(nn.Module):
    def __init__(self, inp, oup, kernel, stride, expand, se_ratio, image_size, bn_mom, bn_eps):
        super().__init__()
        self.id_skip = True
        self.stride = stride
        self.inp, self.oup, self.expand = inp, oup, expand
        mid = inp * expand
        if expand != 1:
            self._expand_conv = Conv2dStaticSamePadding(inp, mid, 1, image_size=image_size, bias=False)
            self._bn0 = nn.BatchNorm2d(mid, momentum=bn_mom, eps=bn_eps)
        self._depthwise_conv = Conv2dStaticSamePadding(mid, mid, kernel, stride=stride, image_size=image_size,
                                                       groups=mid, bias=False)
        self._bn1 = nn.BatchNorm2d(mid, momentum=bn_mom, eps=bn_eps)
        image_size = _out_size(image_size, stride)
        sq = max(1, int(inp * se_ratio))
        self._se_reduce = Conv2dStaticSamePadding(mid, sq, 1, image_size=(1, 1))
        self._se_expand = Conv2dStaticSamePadding(sq, mid, 1, image_size=(1, 1))
        self._project_conv = Conv2dStaticSamePadding(mid, oup, 1, image_size=image_size, bias=False)
        self._bn2 = nn.BatchNorm2d(oup, momentum=bn_mom, eps=bn_eps)
        self._swish = nn.SiLU()

    def forward(self, inputs, drop_connect_rate=None):
        x = inputs
        if self.expand != 1:
            x = self._swish(self._bn0(self._expand_conv(x)))
        x = self._swish(self._bn1(self._depthwise_conv(x)))
        s = F.adaptive_avg_pool2d(x, 1)
        s = self._se_expand(self._swish(self._se_reduce(s)))
        x = torch.sigmoid(s) * x
        x = self._bn2(self._project_conv(x))
        if self.id_skip and self.stride == 1 and self.inp == self.oup:
            if drop_connect_rate and self.training:
                keep = 1 - drop_connect_rate
                mask = torch.floor(keep + torch.rand([x.shape[0], 1, 1, 1], dtype=x.dtype, device=x.device))
                x = x / keep * mask
            x = x + inputs
        return x


class EfficientNetEncoder(nn.Module):
    """smp.encoders.efficientnet.EfficientNetEncoder: 6 feature maps at strides 1,2,4,8,16,32."""

    def __init__(self, model_name="efficientnet-b4", in_channels=3, depth=5, drop_connect_rate=0.2):
        super().__init__()
        width, dcoef, res, _ = _PARAMS[model_name]
        bn_mom, bn_eps = 1 - 0.99, 1e-3
        self._depth = depth
        self._in_channels = in_channels
        self._stage_idxs = _STAGE_IDXS[model_name]
        self._out_channels = (in_channels,) + tuple(_OUT_CHANNELS[model_name][1:])
        self._drop_connect_rate = drop_connect_rate
        image_size = res
        stem = _round_filters(32, width)
        self._conv_stem = Conv2dStaticSamePadding(in_channels, stem, 3, stride=2, image_size=image_size, bias=False)
        self._bn0 = nn.BatchNorm2d(stem, momentum=bn_mom, eps=bn_eps)
        image_size = _out_size(image_size, 2)
        blocks = []
        for (r, k, s, e, i, o, se) in _BASE_BLOCKS:
            i, o, r = _round_filters(i, width), _round_filters(o, width), _round_repeats(r, dcoef)
            blocks.append(MBConvBlock(i, o, k, s, e, se, image_size, bn_mom, bn_eps))
            image_size = _out_size(image_size, s)
            for _ in range(r - 1):
                blocks.append(MBConvBlock(o, o, k, 1, e, se, image_size, bn_mom, bn_eps))
        self._blocks = nn.ModuleList(blocks)
        head_in = _round_filters(320, width)
        head_out = _round_filters(1280, width)
        # kept (and unused in forward) exactly like smp, which only deletes `_fc`
        self._conv_head = Conv2dStaticSamePadding(head_in, head_out, 1, image_size=image_size, bias=False)
        self._bn1 = nn.BatchNorm2d(head_out, momentum=bn_mom, eps=bn_eps)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(_PARAMS[model_name][3])
        self._swish = nn.SiLU()

    @property
    def out_channels(self):
        return self._out_channels[: self._depth + 1]

    def forward(self, x) -> List[torch.Tensor]:
        feats = [x]
        x = self._swish(self._bn0(self._conv_stem(x)))
        feats.append(x)
        bounds = list(self._stage_idxs[: self._depth - 1]) + [len(self._blocks)] if self._depth > 1 else []
        start = 0
        n = len(self._blocks)
        for end in bounds[: max(self._depth - 1, 0)]:
            for idx in range(start, end):
                x = self._blocks[idx](x, self._drop_connect_rate * idx / n)
            feats.append(x)
            start = end
        return feats[: self._depth + 1]


def get_encoder(name, in_channels=3, depth=5, weights=None, output_stride=32, **kwargs):
    """smp.encoders.get_encoder for the EfficientNet family the reference uses
    (train.py:143 passes "efficientnet-b4", in_channels=8, weights left at "imagenet")."""
    if name not in _PARAMS:
        raise KeyError(f"Wrong encoder name `{name}`, supported by this stand-in: {sorted(_PARAMS)}")
    if weights is not None:
        warnings.warn(f"encoder_weights={weights!r}: pretrained ImageNet weights cannot be downloaded in this "
                      "environment; the encoder is randomly initialised (load a checkpoint afterwards)")
    return EfficientNetEncoder(name, in_channels=in_channels, depth=depth)


# ------------------------------------------------------------------ U-Net decoder
class Conv2dReLU(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, use_batchnorm=True):
        conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         bias=not use_batchnorm)
        bn = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        super().__init__(conv, bn, nn.ReLU(inplace=True))


class _Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.attention = nn.Identity()

    def forward(self, x):
        return self.attention(x)


class DecoderBlock(nn.Module):
    def __init__(self, in_channels, skip_channels, out_channels, use_batchnorm=True, attention_type=None):
        super().__init__()
        if attention_type is not None:
            raise NotImplementedError("attention_type is not used by the reference (mymodels.py:251)")
        self.conv1 = Conv2dReLU(in_channels + skip_channels, out_channels, 3, padding=1, use_batchnorm=use_batchnorm)
        self.attention1 = _Attention()
        self.conv2 = Conv2dReLU(out_channels, out_channels, 3, padding=1, use_batchnorm=use_batchnorm)
        self.attention2 = _Attention()

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
            x = self.attention1(x)
        x = self.conv2(self.conv1(x))
        return self.attention2(x)


class UnetDecoder(nn.Module):
    """smp.decoders.unet.decoder.UnetDecoder (0.3.x signature, `use_batchnorm=` kwarg)."""

    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int], n_blocks=5,
                 use_batchnorm=True, attention_type=None, center=False):
        super().__init__()
        if n_blocks != len(decoder_channels):
            raise ValueError(f"Model depth is {n_blocks}, but you provide `decoder_channels` for "
                             f"{len(decoder_channels)} blocks.")
        if center:
            raise NotImplementedError("center block (vgg encoders) is not used by the reference")
        enc = list(encoder_channels[1:])[::-1]
        head = enc[0]
        in_ch = [head] + list(decoder_channels[:-1])
        skip_ch = list(enc[1:]) + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList([DecoderBlock(i, s, o, use_batchnorm, attention_type)
                                     for i, s, o in zip(in_ch, skip_ch, decoder_channels)])
        for m in self.modules():  # smp.base.initialization.initialize_decoder
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, *features):
        features = features[1:][::-1]
        x = self.center(features[0])
        skips = features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x
