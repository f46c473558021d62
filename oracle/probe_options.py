"""TEST INFRASTRUCTURE ONLY -- what the LIVE reference does with the constructor / forward options the sm_100a drop-ins
treat specially (VERDICT r1 "options that raise instead of running").

Runs each option through the reference modules imported from /root/reference (oracle/ref_loader.py) on the CPU and
records whether the reference itself can execute it.  Output: profiles/r02_reference_option_probe.txt (committed), which
DESIGN.md section 1 cites.  Usage: python -m oracle.probe_options
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader, synth  # noqa: E402


def _when2com(**kw):
    ref_loader.install()
    import importlib
    W = importlib.import_module("coperception.models.det.When2com").When2com
    with contextlib.redirect_stdout(io.StringIO()):
        return W(ref_loader.ref_config(), **dict(dict(layer=3, num_agent=5), **kw))


def _run(label, fn, lines):
    try:
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), ref_loader.cpu_cuda_shim(), ref_loader.to_cuda_shim():
            r = fn()
        lines.append("%-58s runs: %s" % (label, r))
    except Exception as e:  # noqa: BLE001  (the point is to record whatever the reference raises)
        tb = traceback.extract_tb(e.__traceback__)[-1]
        where = "%s:%d" % (os.path.basename(tb.filename), tb.lineno)
        lines.append("%-58s RAISES %s at %s: %s" % (label, type(e).__name__, where, str(e).split("\n")[0][:110]))


def main():
    lines = ["# live reference (CPU) on the options the drop-in modules treat specially; made by oracle/probe_options.py"]
    bevs, trans, nat = synth.make_scene(1, 5, 40)
    sd = synth.when2com_det_state(40)

    def w2c(ctor=None, strict=True, **fw):
        m = _when2com(**(ctor or {}))
        m.load_state_dict(sd, strict=strict)
        m.eval()
        r = m(bevs, trans, nat, training=False, batch_size=1, **fw)
        return "cls.sum %.6f" % float(r["cls"].double().sum())

    _run("When2com(sparse=False)  [baseline]", lambda: w2c(), lines)
    _run("When2com(sparse=True)   [attention_net ignores `sparse`]", lambda: w2c(dict(sparse=True)), lines)
    _run("When2com(has_query=False) [query = ones]", lambda: w2c(dict(has_query=False), strict=False), lines)
    _run("When2com.forward(MO_flag=False)", lambda: w2c(MO_flag=False), lines)
    _run("When2com(layer=2, warp_flag=1) inference=activated", lambda: w2c(dict(layer=2)), lines)
    _run("When2com(layer=2) inference=argmax_test", lambda: w2c(dict(layer=2), inference="argmax_test"), lines)
    _run("When2com(layer=4, warp_flag=1) inference=activated", lambda: w2c(dict(layer=4)), lines)

    import importlib
    Config = importlib.import_module("coperception.configs.Config").Config
    F = importlib.import_module("coperception.models.det.FaFNet").FaFNet

    def faf(**cfg_attrs):
        cfg = Config("train", binary=True, only_det=True)
        for k, v in cfg_attrs.items():
            setattr(cfg, k, v)
        m = F(cfg, layer=3, kd_flag=0, num_agent=5)
        m.eval()
        r = m(bevs[:1], batch_size=1)
        return "keys %s" % sorted(r.keys())

    _run("FaFNet(config.use_map=True)", lambda: faf(use_map=True), lines)
    _run("FaFNet(config.use_vis=True)", lambda: faf(use_vis=True), lines)
    _run("FaFNet(config.motion_state=True)", lambda: faf(motion_state=True), lines)

    V = importlib.import_module("coperception.models.det.V2VNet").V2VNet

    def v2v(layer, ch):
        m = V(ref_loader.ref_config(), 3, layer, ch, num_agent=5)
        m.eval()
        r = m(bevs, trans, nat, batch_size=1)
        return "cls %s" % (tuple(r["cls"].shape),)

    _run("V2VNet(layer=4, layer_channel=512)", lambda: v2v(4, 512), lines)
    _run("V2VNet(layer=0, layer_channel=32)", lambda: v2v(0, 32), lines)

    out = os.path.join(ROOT, "profiles", "r02_reference_option_probe.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
