"""bhsr — B200-native hot path of lauraset/Super-resolution-building-height-estimation.

RRDBNet x4 feature extractor + feature-aggregation height head behind the reference's own
nn.Module surface (SR/rrdbnet_arch.py, SR/RRDBNet.py, SR/HRfuse.py, mymodels.py,
aggregate_utils.py), computed by hand-written sm_100a kernels in lib/libbhsr.so (C ABI in
include/bhsr.h).  Import as `bhsr` (alias module at the repo root).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
