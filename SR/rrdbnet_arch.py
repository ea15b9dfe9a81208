"""`from SR.rrdbnet_arch import RealESRGAN` (train.py:14, predict_realesanet_feature_globe.py:16)
resolved to the B200 implementation.  Reference: SR/rrdbnet_arch.py."""
import bhsr  # noqa: F401  (alias of super-resolution-building-height-estimation_b200)
from bhsr.rrdbnet import (RRDB, RRDBNet, RealESRGAN, ResidualDenseBlock, default_init_weights,  # noqa: F401
                          make_layer, pixel_unshuffle)
