"""coperception.models.det.SumFusion on the sm_100a path (reference: CP/models/det/SumFusion.py)."""
from ._fusion import FusionBase


class SumFusion(FusionBase):
    """Sum fusion of the target's map with its warped neighbours ("sum" reduce mode of v2x_warp_reduce_fwd)."""
    KIND = "sum"
