"""coperception.models.det.MaxFusion on the sm_100a path (reference: CP/models/det/MaxFusion.py)."""
from ._fusion import FusionBase


class MaxFusion(FusionBase):
    """Max fusion of the target's map with its warped neighbours ("max" reduce mode of v2x_warp_reduce_fwd)."""
    KIND = "max"
