"""GPU parity of the segmentation path: UNet pieces and the three seg models through
the drop-in ``coperception.models.seg`` modules, vs the CPU oracle and the live-reference fixtures.
Tolerance: 1e-3 relative for the "mixed" (default) and "fp16x3" modes; argmax over classes identical outside the error
margin; bf16 (outside the parity contract) is held to a stated bf16-sized bound."""
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


@pytest.mark.parametrize("mode", ["mixed", "fp16x3", "bf16"])
@pytest.mark.parametrize("tag,kind", CASES, ids=[c[0] for c in CASES])
def test_seg_models(tag, kind, mode, golden_dir, parity_log):
    check_seg_model(tag, kind, mode, golden_dir, parity_log)


def check_seg_model(tag, kind, mode, golden_dir, parity_log):
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
            opts = [int(v) for v in g["options"]] if "options" in g.files else [1, 0]   # (has_query, sparse)
            sd = synth.seg_when2com_state(seed, has_query=bool(opts[0]))
            st = restate.seg_when2com_forward(x, trans, nat, sd, agent_num=a, warp_flag=warp, inference=inference,
                                              stages=True, has_query=bool(opts[0]))
            ref, attn_ref = st["logits"], st["attn"]
            model = When2Com_UNet(default_det_config(), n_classes=8, warp_flag=warp, num_agent=a, has_query=bool(opts[0]),
                                  sparse=bool(opts[1]))
    model.load_state_dict(sd, strict=True)
    model.precision = mode
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
    print("seg %s %s rel_err=%.3e golden=%.3e argmax flips %d (outside margin %d)" % (tag, mode, e, eg, flips, bad))
    rec = dict(logits=e, logits_golden=eg, argmax_flips=flips, argmax_flips_outside_margin=bad)
    gates_agree = True
    if kind == "when2com":
        # discrete gate (p > 0.2): see tests/test_gpu_nets.py::test_when2com_det_forward for the stated bound
        attn = model._plan_list()[0].attn.cpu()
        attn_err = (attn - attn_ref).abs().max().item()
        eye = 0.001 * torch.eye(a).unsqueeze(0)
        differ = ((attn + eye) > 0.2) != ((attn_ref + eye) > 0.2)
        near = ((attn_ref + eye) - 0.2).abs() <= 2 * attn_err
        assert attn_err < (5e-2 if mode == "bf16" else 1e-3), attn_err
        assert not bool((differ & ~near).any())
        gates_agree = not bool(differ.any())
        rec.update(attn_err=attn_err, gates_agree=gates_agree)
    parity_log(tag, mode, **rec)
    if mode != "bf16":
        assert gates_agree and e < 1e-3 and eg < 1e-3 and bad == 0
    elif gates_agree:
        assert e < 8e-2 and eg < 8e-2
