"""coperception.models.det -- every class the reference exports (CP/models/det/__init__.py:1-10) on the sm_100a path:
V2VNet, When2com, FaFNet, TeacherNet, DiscoNet and the Sum / Mean / Max / Cat / AgentWiseWeighted fusion baselines."""
import os as _os

from ... import _extend_with_reference

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]
_ref = _extend_with_reference(__path__, ("models", "det"))

from .V2VNet import V2VNet  # noqa: E402,F401
from .FaFNet import FaFNet  # noqa: E402,F401
from .When2com import When2com  # noqa: E402,F401
from .TeacherNet import TeacherNet  # noqa: E402,F401
from .MeanFusion import MeanFusion  # noqa: E402,F401
from .MaxFusion import MaxFusion  # noqa: E402,F401
from .SumFusion import SumFusion  # noqa: E402,F401
from .CatFusion import CatFusion  # noqa: E402,F401
from .AgentWiseWeightedFusion import AgentWiseWeightedFusion  # noqa: E402,F401
from .DiscoNet import DiscoNet  # noqa: E402,F401
