"""v2x_b200 -- python binding of the sm_100a collaborative-perception kernels (libv2x_b200.so).

``ops``   tensor-level wrappers of the C ABI (include/v2x_b200.h)
``nets``  whole-forward plans (V2VNet det, FaFNet) replayed as CUDA graphs
The drop-in nn.Module surface lives next to this package in ``coperception/``.
"""
from ._lib import LIB_PATH, V2XError, load  # noqa: F401
