"""coperception.models.det.MeanFusion on the sm_100a path (reference: CP/models/det/MeanFusion.py)."""
from ._fusion import FusionBase


class MeanFusion(FusionBase):
    """Mean fusion of the target's map with its warped neighbours ("mean" reduce mode of v2x_warp_reduce_fwd)."""
    KIND = "mean"
