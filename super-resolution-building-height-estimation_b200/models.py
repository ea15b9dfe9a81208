"""Drop-in `SRRegress_Cls_feature` (reference: mymodels.py:233-337 — the only class the two entry
scripts instantiate, train.py:143-148 / predict_realesanet_feature_globe.py:104-108).

Same constructor kwargs, children (`encoder`, `decoder1`, `decoder2`, `reg`, `seg`, `hrfeat`,
`aggre_height`) and forward / forward_unsup / forward_nobuild wiring.  `hrfeat`, `reg`, `seg`
and `aggre_height` run on the libbhsr.so head kernels; the encoder/decoders are the third-party
smp part (see smp_compat.py).  The reference file itself does not import in any Python
(IndentationError at mymodels.py:467); the ablation classes in it are out of scope.
"""
from __future__ import annotations

import torch
from torch import nn

from .hrfuse import HRfeature, HRfuse_residual, conv2d
from .smp_compat import UnetDecoder, get_encoder


class SRRegress_Cls_feature(torch.nn.Module):
    def __init__(self, encoder_name="resnet50", encoder_weights="imagenet", encoder_depth=5,
                 in_channels=7, classes=1, super_in=4, super_mid=64, upscale=4,
                 isaggre=False, chans_build=2, uniform_range=0.3, isunsup=False):
        super().__init__()
        self.encoder = get_encoder(encoder_name, in_channels=in_channels, depth=encoder_depth,
                                   weights=encoder_weights)
        dec_in = (256, 128, 64, 32, 16)
        self.decoder1 = UnetDecoder(encoder_channels=self.encoder.out_channels, decoder_channels=dec_in,
                                    n_blocks=encoder_depth, use_batchnorm=True,
                                    center=True if encoder_name.startswith("vgg") else False,
                                    attention_type=None)
        self.decoder2 = UnetDecoder(encoder_channels=self.encoder.out_channels, decoder_channels=dec_in,
                                    n_blocks=encoder_depth, use_batchnorm=True,
                                    center=True if encoder_name.startswith("vgg") else False,
                                    attention_type=None)
        self.reg = HRfuse_residual(hr_chans=super_mid, lr_chans=dec_in[-1], mid_chans=dec_in[-1],
                                   out_chans=1, upscale=upscale)
        self.seg = HRfuse_residual(hr_chans=super_mid, lr_chans=dec_in[-1], mid_chans=dec_in[-1],
                                   out_chans=chans_build, upscale=upscale)
        self.hrfeat = HRfeature(in_chans=super_in, mid_chans=super_mid, out_chans=super_mid)
        self.isaggre = isaggre
        if self.isaggre:
            self.aggre_height = nn.Conv2d(super_mid, 1, 3, 1, 1)

    def _aggre(self, height_fea):
        return conv2d(height_fea, self.aggre_height.weight, self.aggre_height.bias)

    def forward(self, x, super_fea):
        """(x [B,C,64,64], super_fea [B,64,256,256]) -> height [B,1,256,256], build [B,K,256,256]
        [, height_aggre [B,1,64,64]]   (mymodels.py:270-293)."""
        encode_fea = self.encoder(x)
        super_fea = self.hrfeat(super_fea)
        height_fea = self.decoder1(*encode_fea)
        if self.isaggre:
            height_aggre = self._aggre(height_fea)
        height = self.reg(height_fea, super_fea)
        build = self.decoder2(*encode_fea)
        build = self.seg(build, super_fea)
        if self.isaggre:
            return height, build, height_aggre
        return height, build

    def smp_channels_last(self, enable: bool = True):
        """Run the third-party encoder / decoders in torch.channels_last (NHWC): cuDNN then needs no NCHW<->NHWC
        conversion kernels around its convolutions and takes its own depthwise-conv weight-gradient kernels instead of
        PyTorch's native NCHW one.  Parameters keep their values and state_dict keys; only strides change.  The decoder
        outputs are made NCHW-contiguous again for the head kernels.  Opt-in deployment choice for the stock-PyTorch
        part of the model; affects `forward_smp` (and therefore dp.GraphedTrainStep) only."""
        fmt = torch.channels_last if enable else torch.contiguous_format
        for m in (self.encoder, self.decoder1, self.decoder2):
            m.to(memory_format=fmt)
        self._smp_nhwc = bool(enable)
        return self

    def forward_smp(self, x):
        """The third-party part of `forward` alone: encoder + both U-Net decoders -> (height_fea, build_fea).  It does not
        depend on `super_fea`, so a caller may run it on a second stream next to the frozen RRDBNet forward
        (dp.GraphedTrainStep); `forward_head` is the rest."""
        if getattr(self, "_smp_nhwc", False):
            encode_fea = self.encoder(x.contiguous(memory_format=torch.channels_last))
            return self.decoder1(*encode_fea).contiguous(), self.decoder2(*encode_fea).contiguous()
        encode_fea = self.encoder(x)
        return self.decoder1(*encode_fea), self.decoder2(*encode_fea)

    def forward_head(self, height_fea, build_fea, super_fea):
        """`forward` after `forward_smp`: hrfeat, aggre_height, reg, seg (mymodels.py:276-293, same arithmetic)."""
        super_fea = self.hrfeat(super_fea)
        if self.isaggre:
            height_aggre = self._aggre(height_fea)
        height = self.reg(height_fea, super_fea)
        build = self.seg(build_fea, super_fea)
        if self.isaggre:
            return height, build, height_aggre
        return height, build

    def forward_unsup(self, x, super_fea):
        """mymodels.py:295-312: height only, squeezed."""
        encode_fea = self.encoder(x)
        super_fea = self.hrfeat(super_fea)
        height_fea = self.decoder1(*encode_fea)
        height = self.reg(height_fea, super_fea)
        return height.squeeze()

    def forward_nobuild(self, x, super_fea):
        """mymodels.py:314-337."""
        encode_fea = self.encoder(x)
        super_fea = self.hrfeat(super_fea)
        height_fea = self.decoder1(*encode_fea)
        if self.isaggre:
            height_aggre = self._aggre(height_fea)
        height = self.reg(height_fea, super_fea)
        if self.isaggre:
            return height, height_aggre
        return height
