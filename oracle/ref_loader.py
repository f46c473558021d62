"""TEST INFRASTRUCTURE ONLY -- import the LIVE reference modules from /root/reference.

In the build container the modules come from /root/reference; on the GPU box from the staged copy under
oracle/_ref (oracle/make_ref.py).  Three shims,
none of which touches reference files:
  1. ``import coperception`` fails (CP/__init__.py:4 -> datasets -> shapely), so namespace
     stubs for the packages are pre-seeded in ``sys.modules`` with ``__path__`` pointing into
     the reference tree; leaf modules (models/det/V2VNet.py ...) then import cleanly.
  2. ``collections.Iterable`` alias for CP/utils/convolutional_rnn/utils.py:10.
  3. (seg When2Com_UNet only) ``.cuda()`` neutralised for CPU runs.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib
import os
import sys
import types

# where the reference lives: $V2X_REFERENCE_ROOT, else the read-only tree of the build container, else the copy
# staged by oracle/make_ref.py (git-ignored oracle/_ref, which travels to the GPU box)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("V2X_REFERENCE_ROOT") or (
    "/root/reference/coperception" if os.path.isdir("/root/reference/coperception/coperception/models/det") else _STAGED)
CP = os.path.join(REF_ROOT, "coperception")


def available() -> bool:
    return os.path.isdir(os.path.join(CP, "models", "det"))


def _stub(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def install():
    """Make ``coperception.models.det.X`` importable from the reference tree.  Idempotent."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % CP)
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable  # shim 2
    if getattr(sys.modules.get("coperception"), "_v2x_ref_stub", False):
        return
    for k in [k for k in sys.modules if k == "coperception" or k.startswith("coperception.")]:
        del sys.modules[k]
    root = _stub("coperception", CP)
    root._v2x_ref_stub = True
    for sub in ("models", "utils", "configs", "datasets"):
        _stub("coperception." + sub, os.path.join(CP, sub))
    _stub("coperception.models.det", os.path.join(CP, "models", "det"))
    _stub("coperception.models.seg", os.path.join(CP, "models", "seg"))
    # coperception.models.det.base is a real package whose __init__ only imports leaf modules
    importlib.import_module("coperception.models.det.base")


def uninstall():
    for k in [k for k in sys.modules if k == "coperception" or k.startswith("coperception.")]:
        del sys.modules[k]


def ref_config():
    install()
    Config = importlib.import_module("coperception.configs.Config").Config
    return Config("train", binary=True, only_det=True)


def ref_v2vnet_det(gnn_iter_times=3, layer=3, layer_channel=256, num_agent=5, compress_level=0):
    install()
    V2VNet = importlib.import_module("coperception.models.det.V2VNet").V2VNet
    return V2VNet(ref_config(), gnn_iter_times, layer, layer_channel, num_agent=num_agent,
                  compress_level=compress_level)


def ref_fafnet(num_agent=5, kd_flag=0, compress_level=0):
    install()
    FaFNet = importlib.import_module("coperception.models.det.FaFNet").FaFNet
    return FaFNet(ref_config(), layer=3, kd_flag=kd_flag, num_agent=num_agent, compress_level=compress_level)


def ref_when2com_det(warp_flag=1, num_agent=5, has_query=True, sparse=False, layer=3):
    install()
    When2com = importlib.import_module("coperception.models.det.When2com").When2com
    return When2com(ref_config(), layer=layer, warp_flag=warp_flag, num_agent=num_agent, has_query=has_query, sparse=sparse)


def _seg_config():
    cfg = ref_config()
    return cfg


def ref_seg_unet(n_classes=8, compress_level=0, kd_flag=False):
    install()
    UNet = importlib.import_module("coperception.models.seg.UNet").UNet
    return UNet(13, n_classes, kd_flag=kd_flag, compress_level=compress_level)


def ref_seg_v2vnet(n_classes=8, num_agent=5):
    install()
    V2VNet = importlib.import_module("coperception.models.seg.V2VNet").V2VNet
    return V2VNet(13, n_classes, num_agent=num_agent)


def ref_seg_when2com(n_classes=8, num_agent=5, warp_flag=1, has_query=True, sparse=False):
    install()
    W = importlib.import_module("coperception.models.seg.When2Com_UNet").When2Com_UNet
    return W(_seg_config(), in_channels=13, n_classes=n_classes, warp_flag=warp_flag, num_agent=num_agent,
             has_query=has_query, sparse=sparse)


_FUSION_CLASSES = {"mean": "MeanFusion", "max": "MaxFusion", "sum": "SumFusion", "cat": "CatFusion",
                   "agent": "AgentWiseWeightedFusion", "disco": "DiscoNet"}


def ref_fusion_det(kind, num_agent=5, kd_flag=0, only_v2i=False, compress_level=0, layer=3):
    install()
    name = _FUSION_CLASSES[kind]
    cls = getattr(importlib.import_module("coperception.models.det." + name), name)
    return cls(ref_config(), layer=layer, kd_flag=kd_flag, num_agent=num_agent, only_v2i=only_v2i,
               compress_level=compress_level)


def ref_fusion_seg(kind, n_classes=8, num_agent=5, only_v2i=False):
    install()
    name = _FUSION_CLASSES[kind]
    cls = getattr(importlib.import_module("coperception.models.seg." + name), name)
    if kind == "disco":
        return cls(13, n_classes, num_agent, kd_flag=False, only_v2i=only_v2i)
    return cls(13, n_classes, num_agent, 0, only_v2i)


def ref_teacher():
    install()
    TeacherNet = importlib.import_module("coperception.models.det.TeacherNet").TeacherNet
    return TeacherNet(ref_config())


class cpu_cuda_shim:
    """Shim 3: neutralise the hard-coded ``.cuda()`` of seg When2Com_UNet (When2Com_UNet.py:245) for CPU runs."""

    def __enter__(self):
        import torch
        self._t = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._t
        return False


class to_cuda_shim:
    """``torch.ones(...).to("cuda")`` (det When2com.py:245, the has_query=False branch) becomes a no-op for CPU runs."""

    def __enter__(self):
        import torch
        self._to = torch.Tensor.to
        orig = self._to
        torch.Tensor.to = lambda t, *a, **k: t if (a and isinstance(a[0], str) and a[0] == "cuda") else orig(t, *a, **k)
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.to = self._to
        return False


class float64_shim:
    """Run the live reference modules in float64: ``.to(torch.float)`` (Backbone.py:101) and ``.float()``
    (DetModelBase.py:160) become casts to double while the shim is active, so a ``module.double()`` forward / backward
    is float64 end to end.  Used only to make the training-step fixtures: train-mode BatchNorm gradients of the seeded
    random nets are reproducible to ~1% per element in float32 (cancellation in the BN backward), but to 1e-9 in
    float64, which is what pins the restatement's ALGORITHM exactly."""

    def __enter__(self):
        import torch
        self._to, self._float = torch.Tensor.to, torch.Tensor.float
        orig_to = self._to

        def to(t, *a, **k):
            a = tuple(torch.float64 if x is torch.float32 else x for x in a)
            if k.get("dtype") is torch.float32:
                k["dtype"] = torch.float64
            return orig_to(t, *a, **k)

        torch.Tensor.to = to
        torch.Tensor.float = lambda t, *a, **k: t.double()
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.to, torch.Tensor.float = self._to, self._float
        return False
