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


def test_struct_layout_matches_header(tmp_path):
    """sizeof / field offsets of the ctypes mirror equal what a C compiler makes of include/v2x_b200.h."""
    import subprocess
    from v2x_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "v2x_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(v2x_conv_params), offsetof(v2x_conv_params, weights),'
                   ' offsetof(v2x_conv_params, gru_add), offsetof(v2x_conv_params, tail_cout));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    size, o_w, o_g, o_t = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    P = _lib.ConvParams
    assert (ctypes.sizeof(P), P.weights.offset, P.gru_add.offset, P.tail_cout.offset) == (size, o_w, o_g, o_t)


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


def test_fusion_family_state_dict_schema():
    """Every class of CP/models/{det,seg}/__init__.py exists in the drop-in package with the reference's parameter names
    (strict load of reference-format state dicts; the synthetic dicts are themselves pinned to the live reference by
    tests/test_oracle_golden.py)."""
    import coperception.models.det as det
    import coperception.models.seg as seg
    from oracle import synth
    from v2x_b200 import default_det_config
    names = {"mean": "MeanFusion", "max": "MaxFusion", "sum": "SumFusion", "cat": "CatFusion",
             "agent": "AgentWiseWeightedFusion", "disco": "DiscoNet"}
    for kind, name in names.items():
        m = getattr(det, name)(default_det_config(), kd_flag=0)
        m.load_state_dict(synth.fusion_det_state(kind, 0), strict=True)
        cls = getattr(seg, name)
        m = cls(13, 8, 5, kd_flag=False) if kind == "disco" else cls(13, 8, 5, 0, False)
        m.load_state_dict(synth.seg_fusion_state(kind, 0), strict=True)
    det.TeacherNet(default_det_config()).load_state_dict(synth.fafnet_state(0), strict=True)
    for name in ("V2VNet", "When2com", "FaFNet", "TeacherNet", "DiscoNet", "SumFusion", "MeanFusion", "MaxFusion",
                 "CatFusion", "AgentWiseWeightedFusion"):
        assert hasattr(det, name), name
    for name in ("UNet", "V2VNet", "When2Com_UNet", "MeanFusion", "MaxFusion", "SumFusion", "CatFusion",
                 "AgentWiseWeightedFusion", "DiscoNet", "SegModelBase", "FusionBase"):
        assert hasattr(seg, name), name


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under v2x-sim_b200/ may import it, and bench.py may only touch it in
    its cpu_baseline / --impl reference leg (cpu_reference), never on the measured arm (run_ours)."""
    import ast
    import inspect
    offenders = []
    for d, _, files in os.walk(os.path.join(ROOT, "v2x-sim_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(d, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.level == 0:
                    mods = [node.module or ""]
                if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import textwrap

    def imports_oracle(fn):
        tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
        for node in ast.walk(tree):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else \
                   [node.module or ""] if isinstance(node, ast.ImportFrom) else []
            if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                return True
        return False

    for fn in (bench.run_v2v_det, bench.run_generic, bench.roofline_of, bench.HostPipeline.steps, bench.cpu_baseline_subprocess):
        assert not imports_oracle(fn), fn
    assert imports_oracle(bench._cpu_forwards)   # the one sanctioned use: the CPU baseline / reference arm
