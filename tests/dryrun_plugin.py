"""pytest plugin (opt-in: ``PYTHONPATH=tests python -m pytest -p dryrun_plugin -m gpu tests``): run the GPU test files on a
CPU-only host with every kernel replaced by a no-op.

Purpose: the GPU suite only ever executes on a B200 box.  A change made without GPU time can break the PYTHON of a GPU test
or of the module code it drives -- a keyword that no longer exists, an attribute never set, a ``pytest.raises`` whose
expectation a new feature invalidated -- and nobody sees it until the round-end run stops at it.  Under this plugin all
of that python runs: CUDA tensors live on the CPU, the library is tests/test_plan_dryrun.py::FakeLib (which validates the
arity and types of every ctypes call), streams / events / graphs are dummies.  Numeric assertions then fail, of course
(the outputs are uninitialised memory); tools/gpu_suite_dryrun.py sorts the failures and reports only the ones that are
NOT numeric: exceptions other than AssertionError, and "DID NOT RAISE".

Nothing here is reachable from the product: it is installed by this plugin only.
"""
import contextlib
import ctypes as C
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200"), _HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

_CPU = torch.device("cpu")
_FAKE_CUDA = torch.device("cuda", 0)


def _is_cuda_dev(d):
    if isinstance(d, torch.device):
        return d.type == "cuda"
    if isinstance(d, str):
        return d.startswith("cuda")
    return False


class _Redirect(torch.overrides.TorchFunctionMode):
    """Every torch call that names a CUDA device gets the CPU instead; Tensor.cuda() is the identity."""

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        name = getattr(func, "__name__", "")
        if name == "cuda":
            return args[0]
        if _is_cuda_dev(kwargs.get("device")):
            kwargs["device"] = _CPU
        if name in ("to", "pin_memory") and len(args) >= 2 and _is_cuda_dev(args[1]):
            args = (args[0], _CPU) + tuple(args[2:])
        if name == "pin_memory":
            return args[0]
        return func(*args, **kwargs)


class _Dummy:
    """Stream / Event / CUDAGraph stand-in: every method is a no-op, usable as a context manager."""
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name == "elapsed_time":
            return lambda *a, **k: 1.0
        return lambda *a, **k: None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def pytest_configure(config):
    from test_plan_dryrun import FakeLib
    from v2x_b200 import ops, train
    fake = FakeLib()
    ops.require_gpu = lambda: fake
    ops._stream = lambda: C.c_void_p(0)
    train._stream = lambda: C.c_void_p(0)
    from v2x_b200 import postproc
    for mod in (postproc,):
        if hasattr(mod, "_stream"):
            mod._stream = lambda: C.c_void_p(0)
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.Tensor.device = property(lambda self: _FAKE_CUDA)
    torch.nn.Module.cuda = lambda self, *a, **k: self
    cu = torch.cuda
    cu.is_available = lambda: True
    cu.synchronize = lambda *a, **k: None
    cu.empty_cache = lambda: None
    cu.set_device = lambda *a, **k: None
    cu.current_device = lambda: 0
    cu.device_count = lambda: 1
    cu.current_stream = lambda *a, **k: _Dummy()
    cu.Stream = _Dummy
    cu.Event = _Dummy
    cu.CUDAGraph = _Dummy
    cu.graph = lambda *a, **k: _Dummy()
    cu.stream = lambda *a, **k: contextlib.nullcontext()
    cu.device = lambda *a, **k: contextlib.nullcontext()
    cu.get_device_capability = lambda *a, **k: (10, 0)
    cu.mem_get_info = lambda *a, **k: (1 << 37, 1 << 37)
    class _DataParallel(torch.nn.Module):      # nn.DataParallel with one visible GPU calls module(*inputs) directly
        def __init__(self, module, device_ids=None, output_device=None, dim=0):
            super().__init__()
            self.module = module

        def forward(self, *a, **k):
            return self.module(*a, **k)
    torch.nn.DataParallel = _DataParallel
    import torch.distributed as dist
    _init = dist.init_process_group

    def init_process_group(backend=None, *a, **k):     # NCCL needs GPUs: the collectives run over gloo on the CPU tensors
        k.pop("device_id", None)
        return _init("gloo", *a, **k)
    dist.init_process_group = init_process_group
    mode = _Redirect()
    mode.__enter__()
    config._v2x_redirect = mode
