"""Reference SR/HRfuse.py on the B200 kernels (`from SR.HRfuse import ...`, mymodels.py:13)."""
import bhsr  # noqa: F401
from bhsr.hrfuse import (BasicBlock, GeoNet, HRfeature, HRfuse, HRfuse_residual, HRfuse_x2, HRupsample,  # noqa: F401
                         Refine_residual, Upsampler, conv1x1, conv3x3, default_conv)
