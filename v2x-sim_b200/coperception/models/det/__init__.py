"""coperception.models.det -- V2VNet, When2com and FaFNet run on the sm_100a path; the remaining reference
classes (CP/models/det/__init__.py:1-10) are re-exported from an installed reference when present."""
import os as _os

from ... import _extend_with_reference

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]
_ref = _extend_with_reference(__path__, ("models", "det"))

from .V2VNet import V2VNet  # noqa: E402,F401
from .FaFNet import FaFNet  # noqa: E402,F401
from .When2com import When2com  # noqa: E402,F401

if _ref is not None:  # pragma: no cover - depends on the environment
    for _name in ("DiscoNet", "SumFusion", "MeanFusion", "MaxFusion", "CatFusion",
                  "AgentWiseWeightedFusion", "TeacherNet"):
        try:
            _mod = __import__(__name__ + "." + _name, fromlist=[_name])
            globals()[_name] = getattr(_mod, _name)
        except Exception:  # reference class not importable here (missing third-party deps)
            pass
