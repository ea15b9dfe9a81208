"""`from aggregate_utils import aggregate_torch` (BH_loader.py:9) resolved to the B200 package."""
import bhsr  # noqa: F401
from bhsr.aggregate import aggregate, aggregate_torch, aggregate_torch_gpu  # noqa: F401
