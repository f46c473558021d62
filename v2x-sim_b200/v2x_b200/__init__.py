"""v2x_b200 -- python binding of the sm_100a collaborative-perception kernels (libv2x_b200.so).

``ops``   tensor-level wrappers of the C ABI (include/v2x_b200.h)
``nets``  whole-forward plans (V2VNet det, FaFNet) replayed as CUDA graphs
The drop-in nn.Module surface lives next to this package in ``coperception/``.
"""
from ._lib import LIB_PATH, V2XError, load  # noqa: F401


def default_det_config():
    """The detection hyper-parameters the models read, with the reference's literal defaults
    (CP/configs/Config.py:95-131,154-177: binary two-class, 6 anchors, 6-value box code, 256x256x13 map).
    For callers that do not have the reference ``Config`` class importable (benchmarks, tests)."""
    import math
    from types import SimpleNamespace
    return SimpleNamespace(
        binary=True, only_det=True, motion_state=False, use_map=False, use_vis=False, pred_len=1,
        box_code_size=6, category_num=2, map_dims=[256, 256, 13],
        anchor_size=[[2.0, 4.0, 0.0], [2.0, 4.0, math.pi / 2.0], [2.0, 4.0, -math.pi / 4.0],
                     [3.0, 12.0, 0.0], [3.0, 12.0, math.pi / 2.0], [3.0, 12.0, -math.pi / 4.0]])
