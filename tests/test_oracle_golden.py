"""CPU: the oracle restatement (oracle/restate.py) against fixtures made from the LIVE reference
modules (oracle/gen_golden.py -> tests/golden/*.npz).  Tolerance 2e-4 absolute on O(1) activations
(the restatement uses functional ops instead of nn.Modules; measured difference 1.5e-5)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, synth
from oracle.gen_golden import STRIDE, argmax_checksum

ATOL = 2e-4


def _check(name, t, g):
    t = t.detach().float().contiguous()
    assert list(t.shape) == list(g[name + ".shape"]), name
    sub = t.view(-1)[::STRIDE].numpy()
    ref = g[name + ".sub"]
    err = np.abs(sub - ref).max()
    assert err < ATOL, (name, err)
    assert abs(t.double().abs().sum().item() - float(g[name + ".abssum"])) < 1e-5 * float(g[name + ".abssum"]) + 1e-3


def test_warp_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_small_seed3.npz"))
    local = torch.from_numpy(g["local"])
    trans = torch.from_numpy(g["trans"])
    C = local.shape[2]
    for k, (b, j, i) in enumerate(g["pairs"]):
        out = restate.feature_transformation(local, int(b), int(j), int(i), trans, (1, C, 32, 32))
        assert np.abs(out.numpy() - g["out"][k]).max() < 1e-5


def test_convgru_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "convgru_small_seed4.npz"))
    sd = {"convgru." + k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    x = torch.from_numpy(g["x"])[0]
    y = restate.convgru_zero_hidden(x, sd)
    assert np.abs(y.numpy() - g["y"][0]).max() < 1e-5


def test_fafnet(golden_dir):
    g = np.load(os.path.join(golden_dir, "fafnet_n2_seed0.npz"))
    n, seed = [int(v) for v in g["meta"]]
    sd = synth.fafnet_state(seed)
    with torch.no_grad():
        r = restate.fafnet_forward(synth.make_bevs(n, seed), sd)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


def test_v2vnet_stages(golden_dir):
    g = np.load(os.path.join(golden_dir, "v2vnet_det_stages_seed0.npz"))
    _, a, seed = [int(v) for v in g["meta"]]
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(1, a, seed)
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, stages=True)
    for i in range(5):
        _check("enc%d" % i, r["enc"][i], g)
    _check("fused", r["fused"], g)
    _check("x8", r["x8"], g)


@pytest.mark.parametrize("tag", ["v2vnet_det_A5B1_seed0", "v2vnet_det_A5B2_seed1_present53"])
def test_v2vnet_det(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, gnn = [int(v) for v in g["meta"]]
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, gnn_iter=gnn)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    cnt, chk = argmax_checksum(r["cls"])
    # argmax can legitimately flip where the two logits agree to ~1e-5; allow a handful
    assert np.abs(cnt - g["cls.argmax_count"]).max() <= 4


@pytest.mark.parametrize("tag", ["when2com_det_warp_activated_seed2", "when2com_det_nowarp_argmax_seed3_present4",
                                 "when2com_det_warp_softmax_B2_seed4"])
def test_when2com_det(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    inference = str(g["inference"])
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    sd = synth.when2com_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        r = restate.when2com_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, warp_flag=warp,
                                         inference=inference)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


@pytest.mark.parametrize("tag,kind", [("seg_unet_seed0", "unet"), ("seg_v2vnet_seed1_present4", "v2vnet"),
                                      ("seg_when2com_warp_activated_seed2", "when2com"),
                                      ("seg_when2com_nowarp_activated_seed3", "when2com")])
def test_seg_models(golden_dir, tag, kind):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    x, trans, nat = synth.make_seg_scene(batch, a, seed, present=present)
    with torch.no_grad():
        if kind == "unet":
            r = restate.seg_unet_forward(x, synth.seg_unet_state(seed))
        elif kind == "v2vnet":
            r = restate.seg_v2vnet_forward(x, trans, nat, synth.seg_v2vnet_state(seed), agent_num=a)
        else:
            r = restate.seg_when2com_forward(x, trans, nat, synth.seg_when2com_state(seed), agent_num=a, warp_flag=warp,
                                             inference=str(g["inference"]))
    _check("logits", r, g)
