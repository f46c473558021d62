"""CPU: the C-ABI library loads and exports every symbol include/v2x_b200.h declares; argument
validation and the no-GPU failure mode are loud (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "v2x_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2x_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from v2x_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 9
    bound = {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), name
        assert name in bound, "python binding missing for " + name
    assert lib.v2x_version() >= 1


def test_struct_layout_matches_header():
    from v2x_b200 import _lib
    # 2 ptr + 2 i32 + 6 i32 + 2 ptr + 6 i32 + 2 ptr + 3 i32 (+pad) + 3 ptr + 2 i32 + 4 i32
    assert ctypes.sizeof(_lib.ConvParams) == 176


def test_argument_validation_is_loud():
    from v2x_b200 import _lib
    lib = _lib.load()
    p = _lib.ConvParams()
    rc = lib.v2x_conv_fwd(ctypes.byref(p), None)
    assert rc == -1 and b"null" in lib.v2x_last_error()
    assert lib.v2x_pack_input(None, None, 0, 13, 16, 1, None) == -1


def test_no_gpu_means_error_not_fallback():
    import torch
    from v2x_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_import_error()):
        ops.require_gpu()


def _import_error():
    from v2x_b200 import V2XError
    return V2XError


def test_dropin_state_dict_schema():
    """The drop-in modules expose exactly the reference's parameter / buffer names and shapes."""
    from coperception.models.det import FaFNet, V2VNet
    from oracle import synth
    from v2x_b200 import default_det_config
    m = V2VNet(default_det_config(), 3, 3, 256)
    want = synth.v2vnet_det_state(0)
    have = m.state_dict()
    assert set(have) == set(want)
    for k in want:
        assert tuple(have[k].shape) == tuple(want[k].shape), k
    f = FaFNet(default_det_config(), kd_flag=0)
    assert set(f.state_dict()) == set(synth.fafnet_state(0))
    # reference checkpoints of det models carry the DataParallel "module." prefix (train_codet.py:261)
    import torch
    dp = torch.nn.DataParallel(m)
    dp.load_state_dict({"module." + k: v for k, v in want.items()}, strict=True)
