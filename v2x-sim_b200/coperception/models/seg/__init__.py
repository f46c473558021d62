"""coperception.models.seg -- every class the reference exports (CP/models/seg/__init__.py:1-11) on the sm_100a path."""
import os as _os

from ... import _extend_with_reference

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]
_ref = _extend_with_reference(__path__, ("models", "seg"))

from .SegModelBase import SegModelBase  # noqa: E402,F401
from .UNet import UNet  # noqa: E402,F401
from .V2VNet import V2VNet  # noqa: E402,F401
from .When2Com_UNet import When2Com_UNet  # noqa: E402,F401
from .FusionBase import FusionBase  # noqa: E402,F401
from .MeanFusion import MeanFusion  # noqa: E402,F401
from .MaxFusion import MaxFusion  # noqa: E402,F401
from .SumFusion import SumFusion  # noqa: E402,F401
from .CatFusion import CatFusion  # noqa: E402,F401
from .AgentWiseWeightedFusion import AgentWiseWeightedFusion  # noqa: E402,F401
from .DiscoNet import DiscoNet  # noqa: E402,F401
