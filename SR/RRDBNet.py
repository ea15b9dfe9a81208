"""Reference SR/RRDBNet.py (ESRGAN-era attribute names) on the B200 kernels."""
import bhsr  # noqa: F401
from bhsr.rrdbnet import OldRRDBNet as RRDBNet  # noqa: F401
from bhsr.rrdbnet import _OldRDB as ResidualDenseBlock_5C  # noqa: F401
from bhsr.rrdbnet import _OldRRDB as RRDB  # noqa: F401
