"""`from mymodels import SRRegress_Cls_feature` (train.py:16) resolved to the B200 implementation.
The reference mymodels.py does not parse (IndentationError at line 467); only its production
class is provided — the ablation zoo is out of scope."""
import bhsr  # noqa: F401
from bhsr.models import SRRegress_Cls_feature  # noqa: F401
