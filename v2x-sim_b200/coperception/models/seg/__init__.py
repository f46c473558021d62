"""coperception.models.seg -- UNet, V2VNet and When2Com_UNet run on the sm_100a path; the other reference classes
(CP/models/seg/__init__.py:1-11) are re-exported from an installed reference when present."""
import os as _os

from ... import _extend_with_reference

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]
_ref = _extend_with_reference(__path__, ("models", "seg"))

from .SegModelBase import SegModelBase  # noqa: E402,F401
from .UNet import UNet  # noqa: E402,F401
from .V2VNet import V2VNet  # noqa: E402,F401
from .When2Com_UNet import When2Com_UNet  # noqa: E402,F401

if _ref is not None:  # pragma: no cover - depends on the environment
    for _name in ("FusionBase", "MeanFusion", "MaxFusion", "SumFusion", "CatFusion", "AgentWiseWeightedFusion",
                  "DiscoNet"):
        try:
            _mod = __import__(__name__ + "." + _name, fromlist=[_name])
            globals()[_name] = getattr(_mod, _name)
        except Exception:
            pass
