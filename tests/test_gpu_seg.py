"""GPU parity of the segmentation path: UNet pieces and the three seg models through
the drop-in ``coperception.models.seg`` modules, vs the CPU oracle and the live-reference fixtures.
Tolerance: 1e-3 relative for bf16x3; argmax over classes identical outside the error margin."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def to_act(x_nchw, planes):
    from v2x_b200 import ops
    return ops.pack_input_nchw(x_nchw.cuda().contiguous(), x_nchw.shape[1], planes)


def test_unet_pieces():
    """NCHW pack, MaxPool2d(2), bilinear x2 (align_corners=True) vs torch."""
    from v2x_b200 import ops
    ops.require_gpu()
    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 16, 32, 48), generator=g)
    act = to_act(x, 2)
    assert rel_err(ops.act_to_float(act), x) < 1e-4
    assert rel_err(ops.act_to_float(ops.maxpool2(act)), F.max_pool2d(x, 2)) < 1e-4
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    assert rel_err(ops.act_to_float(ops.upsample_bilinear2(act)), up) < 1e-4
    x13 = (torch.rand((1, 13, 16, 16), generator=g) < 0.1).float()
    a13 = ops.act_to_float(ops.pack_input_nchw(x13.cuda(), 16, 1)).cpu()
    assert torch.equal(a13[:, :13], x13) and a13[:, 13:].abs().max().item() == 0.0


def _margin_flips(out, ref):
    out, ref = out.detach().float().cpu(), ref.detach().float().cpu()
    err = (out - ref).abs().max().item()
    flip = out.argmax(1) != ref.argmax(1)
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    return int(flip.sum()), int((flip & (margin > 2 * err)).sum())


CASES = [("seg_unet_seed0", "unet"), ("seg_v2vnet_seed1_present4", "v2vnet"),
         ("seg_when2com_warp_activated_seed2", "when2com"), ("seg_when2com_nowarp_activated_seed3", "when2com")]


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("tag,kind", CASES, ids=[c[0] for c in CASES])
def test_seg_models(tag, kind, planes, golden_dir):
    from coperception.models.seg import UNet, V2VNet, When2Com_UNet
    from oracle import restate, synth
    from oracle.gen_golden import STRIDE
    from v2x_b200 import default_det_config
    path = os.path.join(golden_dir, tag + ".npz")
    g = np.load(path)
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    inference = str(g["inference"])
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    x, trans, nat = synth.make_seg_scene(batch, a, seed, present=present)
    with torch.no_grad():
        if kind == "unet":
            sd = synth.seg_unet_state(seed)
            ref = restate.seg_unet_forward(x, sd)
            model = UNet(13, 8)
        elif kind == "v2vnet":
            sd = synth.seg_v2vnet_state(seed)
            ref = restate.seg_v2vnet_forward(x, trans, nat, sd, agent_num=a)
            model = V2VNet(13, 8, num_agent=a)
        else:
            sd = synth.seg_when2com_state(seed)
            ref = restate.seg_when2com_forward(x, trans, nat, sd, agent_num=a, warp_flag=warp, inference=inference)
            model = When2Com_UNet(default_det_config(), n_classes=8, warp_flag=warp, num_agent=a)
    model.load_state_dict(sd, strict=True)
    model.precision = "bf16x3" if planes == 2 else "bf16"
    model = model.cuda().eval()
    with torch.no_grad():
        if kind == "unet":
            out = model(x.cuda())
        elif kind == "v2vnet":
            out = model(x.cuda(), trans.cuda(), nat.cuda())
        else:
            out = model(x.cuda(), trans.cuda(), nat.cuda(), inference=inference, training=False)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    e = rel_err(out, ref)
    sub = out.detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
    eg = float(np.abs(sub - g["logits.sub"]).max() / np.abs(g["logits.sub"]).max())
    flips, bad = _margin_flips(out, ref)
    print("seg %s planes=%d rel_err=%.3e golden=%.3e argmax flips %d (outside margin %d)" % (tag, planes, e, eg, flips, bad))
    if planes == 2:
        assert e < 1e-3 and eg < 1e-3 and bad == 0
    elif kind in ("unet", "v2vnet"):
        assert e < 8e-2
