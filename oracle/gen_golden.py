"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the LIVE reference modules.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The reference has no golden vectors of its own, so these fixtures --
outputs of the unmodified reference modules on seeded synthetic inputs/weights from
``oracle/synth.py`` -- are what pins the CPU restatement (``oracle/restate.py``), which in
turn is what the CUDA path is compared against on the GPU box.

Full outputs are tens of MB, so each fixture stores a strided subsample of every output
tensor plus whole-tensor statistics (sum, abs-sum, argmax histogram / checksum).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_loader, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STRIDE = 389  # prime, so the subsample walks every channel / anchor / row phase


def summarize(name, t, out):
    t = t.detach().to(torch.float32).contiguous()
    flat = t.view(-1)
    out[name + ".shape"] = np.asarray(t.shape, dtype=np.int64)
    out[name + ".sub"] = flat[::STRIDE].numpy().copy()
    out[name + ".sum"] = np.float64(flat.double().sum().item())
    out[name + ".abssum"] = np.float64(flat.double().abs().sum().item())


def argmax_checksum(cls):
    """cls: [N, P, 2] -> per-map count of argmax==1 and a position-weighted checksum (int64)."""
    am = cls.argmax(-1).to(torch.int64)
    idx = torch.arange(am.shape[1], dtype=torch.int64) % 65521
    return am.sum(1).numpy(), (am * idx).sum(1).numpy()


def gen_v2vnet(tag, batch, seed, present=None, gnn_iter=3):
    m = ref_loader.ref_v2vnet_det(gnn_iter_times=gnn_iter)
    sd = synth.v2vnet_det_state(seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=batch)
    out = {"meta": np.asarray([batch, 5, seed, gnn_iter], dtype=np.int64)}
    if present is not None:
        out["present"] = np.asarray(present, dtype=np.int64)
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    out["cls.argmax_count"], out["cls.argmax_checksum"] = argmax_checksum(r["cls"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "loc.sum", out["loc.sum"], "cls.sum", out["cls.sum"])


def gen_when2com(tag, batch, seed, warp_flag, inference, present=None, has_query=True, sparse=False, layer=3):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_loader.ref_when2com_det(warp_flag=warp_flag, has_query=has_query, sparse=sparse, layer=layer)
    sd = synth.when2com_det_state(seed, has_query=has_query)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(batch, 5, seed, present=present)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), ref_loader.to_cuda_shim():
        r = m(bevs, trans, nat, training=False, MO_flag=True, inference=inference, batch_size=batch)
    out = {"meta": np.asarray([batch, 5, seed, warp_flag], dtype=np.int64), "inference": np.asarray(inference),
           "options": np.asarray([int(has_query), int(sparse), layer], dtype=np.int64)}
    if present is not None:
        out["present"] = np.asarray(present, dtype=np.int64)
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    out["cls.argmax_count"], out["cls.argmax_checksum"] = argmax_checksum(r["cls"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "loc.sum", out["loc.sum"], "cls.sum", out["cls.sum"])


def gen_seg(tag, kind, batch, seed, present=None, inference="activated", warp_flag=1, has_query=True, sparse=False):
    """kind in {"unet", "v2vnet", "when2com"}; logits [N,8,256,256] of the live seg models."""
    import contextlib
    import io
    a = 5
    x, trans, nat = synth.make_seg_scene(batch, a, seed, present=present)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "unet":
            m, sd = ref_loader.ref_seg_unet(), synth.seg_unet_state(seed)
        elif kind == "v2vnet":
            m, sd = ref_loader.ref_seg_v2vnet(num_agent=a), synth.seg_v2vnet_state(seed)
        else:
            m = ref_loader.ref_seg_when2com(num_agent=a, warp_flag=warp_flag, has_query=has_query, sparse=sparse)
            sd = synth.seg_when2com_state(seed, has_query=has_query)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad(), ref_loader.cpu_cuda_shim(), ref_loader.to_cuda_shim(), contextlib.redirect_stdout(io.StringIO()):
        if kind == "unet":
            r = m(x)
        elif kind == "v2vnet":
            r = m(x, trans, nat)
        else:
            r = m(x, trans, nat, inference=inference, training=False)
    out = {"meta": np.asarray([batch, a, seed, warp_flag], dtype=np.int64), "inference": np.asarray(inference),
           "kind": np.asarray(kind)}
    if kind == "when2com" and (not has_query or sparse):
        out["options"] = np.asarray([int(has_query), int(sparse)], dtype=np.int64)
    if present is not None:
        out["present"] = np.asarray(present, dtype=np.int64)
    summarize("logits", r, out)
    am = r.argmax(1).to(torch.int64)
    out["logits.argmax_hist"] = np.stack([(am == c).sum((1, 2)).numpy() for c in range(r.shape[1])], 1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "logits.sum", out["logits.sum"])


def gen_fafnet(tag, n, seed):
    m = ref_loader.ref_fafnet(kd_flag=0)
    sd = synth.fafnet_state(seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs = synth.make_bevs(n, seed)
    with torch.no_grad():
        r = m(bevs)
    out = {"meta": np.asarray([n, seed], dtype=np.int64)}
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    out["cls.argmax_count"], out["cls.argmax_checksum"] = argmax_checksum(r["cls"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "loc.sum", out["loc.sum"])


def gen_stages(tag, seed):
    """Encoder layers, fused map and decoder output of the live V2VNet via forward hooks."""
    m = ref_loader.ref_v2vnet_det()
    sd = synth.v2vnet_det_state(seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, seed)
    cap = {}
    m.u_encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("enc", [t.clone() for t in o]))
    m.decoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("dec", (i[3].clone(), o[0].clone())))
    with torch.no_grad():
        m(bevs, trans, nat, batch_size=1)
    out = {"meta": np.asarray([1, 5, seed], dtype=np.int64)}
    for i, t in enumerate(cap["enc"]):
        summarize("enc%d" % i, t, out)
    summarize("fused", cap["dec"][0], out)
    summarize("x8", cap["dec"][1], out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "fused.sum", out["fused.sum"])


def gen_warp(tag, seed):
    """Full outputs of DetModelBase.feature_transformation on a small-channel map."""
    ref_loader.install()
    from coperception.models.det.base.DetModelBase import DetModelBase
    g = torch.Generator().manual_seed(seed)
    C = 8
    local = torch.randn((2, 5, C, 32, 32), generator=g)
    trans = synth.make_trans_matrices(2, 5, seed)
    pairs = [(0, 1, 0), (0, 4, 2), (1, 0, 3), (1, 2, 2), (1, 3, 4)]  # (b, j, i)
    outs = []
    with torch.no_grad():
        for b, j, i in pairs:
            outs.append(DetModelBase.feature_transformation(b, j, i, local, None, None, (1, C, 32, 32), trans))
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), local=local.numpy(), trans=trans.numpy(),
                        pairs=np.asarray(pairs, dtype=np.int64), out=torch.stack(outs).numpy())
    print(tag, "ok")


def gen_convgru(tag, seed):
    """Full output of the reference Conv2dGRU (zero hidden) on a small map."""
    ref_loader.install()
    import coperception.utils.convolutional_rnn as convrnn
    C = 16
    gru = convrnn.Conv2dGRU(in_channels=2 * C, out_channels=C, kernel_size=3, num_layers=1,
                            bidirectional=False, dilation=1, stride=1)
    g = torch.Generator().manual_seed(seed)
    sd = {k: (torch.rand(v.shape, generator=g) - 0.5) * 0.5 for k, v in gru.state_dict().items()}
    gru.load_state_dict(sd)
    x = torch.randn((1, 1, 2 * C, 8, 8), generator=g)
    with torch.no_grad():
        y, _ = gru(x, None)
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), x=x.numpy(), y=y.numpy(),
                        **{"sd." + k: v.numpy() for k, v in sd.items()})
    print(tag, "ok")


def gen_fusion_det(tag, kind, batch, seed, present=None, only_v2i=False):
    """Result dict of the live det FusionBase-family model ``kind`` (kd_flag = 0)."""
    m = ref_loader.ref_fusion_det(kind, only_v2i=only_v2i)
    sd = synth.fusion_det_state(kind, seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=batch)
    if kind == "disco":
        r = r[0]   # (result, save_agent_weight_list) when kd_flag == 0 (DiscoNet.py:125-129)
    out = {"meta": np.asarray([batch, 5, seed, int(only_v2i)], dtype=np.int64), "kind": np.asarray(kind)}
    if present is not None:
        out["present"] = np.asarray(present, dtype=np.int64)
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    out["cls.argmax_count"], out["cls.argmax_checksum"] = argmax_checksum(r["cls"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "loc.sum", out["loc.sum"], "cls.sum", out["cls.sum"])


def gen_fusion_seg(tag, kind, batch, seed, present=None, only_v2i=False):
    m = ref_loader.ref_fusion_seg(kind, only_v2i=only_v2i)
    sd = synth.seg_fusion_state(kind, seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x, trans, nat = synth.make_seg_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        r = m(x, trans, nat)
    out = {"meta": np.asarray([batch, 5, seed, int(only_v2i)], dtype=np.int64), "kind": np.asarray(kind)}
    if present is not None:
        out["present"] = np.asarray(present, dtype=np.int64)
    summarize("logits", r, out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "logits.sum", out["logits.sum"])


def gen_teacher(tag, n, seed):
    m = ref_loader.ref_teacher()
    sd = synth.fafnet_state(seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        r = m(synth.make_bevs(n, seed))
    out = {"meta": np.asarray([n, seed], dtype=np.int64)}
    for name, t in zip(("x8", "x7", "x6", "x5", "x3", "x4"), r):
        summarize(name, t, out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "x8.sum", out["x8.sum"])


FUSION_FIXTURES = [  # (tag, family, kind, batch, seed, present, only_v2i)
    ("fusion_det_mean_seed5", "det", "mean", 1, 5, None, False),
    ("fusion_det_max_seed6_present4", "det", "max", 1, 6, [4], False),
    ("fusion_det_sum_seed7_v2i", "det", "sum", 1, 7, None, True),
    ("fusion_det_cat_B2_seed8_present35", "det", "cat", 2, 8, [3, 5], False),
    ("fusion_det_agent_seed9_present4", "det", "agent", 1, 9, [4], False),
    ("fusion_det_disco_B2_seed10_present53", "det", "disco", 2, 10, [5, 3], False),
    ("fusion_seg_mean_seed11_present4", "seg", "mean", 1, 11, [4], False),
    ("fusion_seg_cat_seed12", "seg", "cat", 1, 12, None, False),
    ("fusion_seg_agent_seed13", "seg", "agent", 1, 13, None, False),
    ("fusion_seg_disco_seed14_present3", "seg", "disco", 1, 14, [3], False),
]


def gen_fusion_all():
    for tag, family, kind, batch, seed, present, v2i in FUSION_FIXTURES:
        (gen_fusion_det if family == "det" else gen_fusion_seg)(tag, kind, batch, seed, present=present, only_v2i=v2i)
    gen_teacher("teacher_n1_seed3", 1, 3)


def gen_compress():
    """compress_level > 0 (Backbone.py:74-87,138-141; SegModelBase.py:29-43): V2VNet det at level 2 (64 channels),
    DiscoNet det at level 6 (4 channels: exercises the zero-padded narrow operand), seg UNet at level 3 (64 of 512)."""
    # V2VNet det
    m = ref_loader.ref_v2vnet_det(compress_level=2)
    sd = synth.v2vnet_det_state(15, compress_level=2)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 15, present=[4])
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=1)
    out = {"meta": np.asarray([1, 5, 15, 2], dtype=np.int64), "present": np.asarray([4], dtype=np.int64)}
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "compress_v2vnet_det_l2_seed15.npz"), **out)
    print("compress_v2vnet_det_l2_seed15", out["loc.sum"])
    # DiscoNet det, 4 compressed channels
    m = ref_loader.ref_fusion_det("disco", compress_level=6)
    sd = synth.fusion_det_state("disco", 16, compress_level=6)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 16)
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=1)[0]
    out = {"meta": np.asarray([1, 5, 16, 6], dtype=np.int64)}
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "compress_disco_det_l6_seed16.npz"), **out)
    print("compress_disco_det_l6_seed16", out["loc.sum"])
    # seg UNet with kd outputs
    m = ref_loader.ref_seg_unet(compress_level=3, kd_flag=True)
    sd = synth.seg_unet_state(17, compress_level=3)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x, _, _ = synth.make_seg_scene(1, 2, 17)
    with torch.no_grad():
        r = m(x)
    out = {"meta": np.asarray([1, 2, 17, 3], dtype=np.int64)}
    for name, t in zip(("logits", "x9", "x8", "x7", "x6", "x5", "x4"), r):
        summarize(name, t, out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "compress_seg_unet_l3_kd_seed17.npz"), **out)
    print("compress_seg_unet_l3_kd_seed17", out["logits.sum"])


def gen_layers():
    """Communication at other encoder layers (DetModelBase.py:71-92): V2VNet at layer 2 (128 ch, 64x64, two GNN
    rounds), DiscoNet at layer 2, MaxFusion at layer 1 (64 ch, 128x128)."""
    def save(tag, r, meta):
        out = {"meta": np.asarray(meta, dtype=np.int64)}
        summarize("loc", r["loc"], out)
        summarize("cls", r["cls"], out)
        np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
        print(tag, out["loc.sum"])
    m = ref_loader.ref_v2vnet_det(gnn_iter_times=2, layer=2, layer_channel=128)
    sd = synth.v2vnet_det_state(18, layer_channel=128)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 18, present=[4])
    with torch.no_grad():
        save("layer2_v2vnet_det_seed18", m(bevs, trans, nat, batch_size=1), [1, 5, 18, 2])
    m = ref_loader.ref_fusion_det("disco", layer=2)
    sd = synth.fusion_det_state("disco", 19, channel=128)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 19)
    with torch.no_grad():
        save("layer2_disco_det_seed19", m(bevs, trans, nat, batch_size=1)[0], [1, 5, 19, 2])
    m = ref_loader.ref_fusion_det("max", layer=1)
    sd = synth.fusion_det_state("max", 20)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 20, present=[3])
    with torch.no_grad():
        save("layer1_max_det_seed20", m(bevs, trans, nat, batch_size=1), [1, 5, 20, 1])


def grad_stride(numel):
    """Subsample stride of a gradient tensor in the training fixtures: about 512 samples of the big ones."""
    return max(1, numel // 512) | 1


TRAIN_CASES = {   # tag -> (kind, seed)
    "train_step_v2vnet_seed21": ("v2vnet", 21),
    "train_step_fafnet_seed22": ("fafnet", 22),
    "train_step_when2com_seed23": ("when2com", 23),
    "train_step_disco_seed24": ("disco", 24),
    "train_step_seg_unet_seed25": ("seg_unet", 25),
    "train_step_seg_v2vnet_seed26": ("seg_v2vnet", 26),
    "train_step_cat_seed27": ("cat", 27),
    "train_step_agent_seed28": ("agent", 28),
    "train_step_seg_when2com_seed29": ("seg_when2com", 29),
    "train_step_mean_seed32": ("mean", 32),
    "train_step_sum_seed33": ("sum", 33),
    "train_step_max_seed34": ("max", 34),
    "train_step_seg_mean_seed35": ("seg_mean", 35),
    "train_step_seg_max_seed36": ("seg_max", 36),
    "train_step_seg_cat_seed37": ("seg_cat", 37),
    "train_step_seg_agent_seed38": ("seg_agent", 38),
    "train_step_seg_disco_seed39": ("seg_disco", 39),
    # compress_level > 0 in .train() (Backbone.py:138-141 / UNet.py:30-32 with train-mode BatchNorm on the pair)
    "train_step_v2vnet_c2_seed45": ("v2vnet_c2", 45),
    "train_step_seg_unet_c3_seed46": ("seg_unet_c3", 46),
}


def train_case(kind, seed):
    """(state_dict, inputs tuple, output keys) of a training-step case; shared by the generator and the CPU test."""
    if kind == "v2vnet":
        return synth.v2vnet_det_state(seed), synth.make_scene(1, 5, seed, present=[4]), ("loc", "cls")
    if kind == "v2vnet_c2":
        return synth.v2vnet_det_state(seed, compress_level=2), synth.make_scene(1, 5, seed, present=[4]), ("loc", "cls")
    if kind == "seg_unet_c3":
        return synth.seg_unet_state(seed, compress_level=3), (synth.make_seg_scene(1, 2, seed)[0],), ("logits",)
    if kind == "fafnet":
        return synth.fafnet_state(seed), (synth.make_bevs(3, seed),), ("loc", "cls")
    if kind == "when2com":
        return synth.when2com_det_state(seed), synth.make_scene(1, 5, seed), ("loc", "cls")
    if kind in ("disco", "cat", "agent", "mean", "sum", "max"):
        return synth.fusion_det_state(kind, seed), synth.make_scene(1, 5, seed, present=[4]), ("loc", "cls")
    if kind == "seg_when2com":
        return synth.seg_when2com_state(seed), synth.make_seg_scene(1, 5, seed), ("logits",)
    if kind == "seg_unet":
        return synth.seg_unet_state(seed), (synth.make_seg_scene(1, 2, seed)[0],), ("logits",)
    if kind == "seg_v2vnet":
        return synth.seg_v2vnet_state(seed), synth.make_seg_scene(1, 5, seed, present=[4]), ("logits",)
    if kind in ("seg_mean", "seg_sum", "seg_max", "seg_cat", "seg_agent", "seg_disco"):
        return synth.seg_fusion_state(kind[4:], seed), synth.make_seg_scene(1, 5, seed, present=[4]), ("logits",)
    raise ValueError(kind)


def make_upstream(shapes, seed):
    """Seeded d(loss)/d(out) per output key: dense small values, like the gradient of a mean-reduced loss (float64)."""
    g = torch.Generator().manual_seed(4000 + seed)
    return {k: (torch.rand(tuple(shapes[k]), generator=g, dtype=torch.float64) - 0.5) * (2.0 / float(np.prod(shapes[k])) ** 0.5)
            for k in sorted(shapes)}


def gen_train_step(tag, kind, seed):
    """One training step of the LIVE reference module in .train() mode, in float64 (ref_loader.float64_shim): outputs,
    the gradient of EVERY parameter that receives one (strided subsample + L2 norm) and every BatchNorm's running
    buffers after the step.  SURVEY 8(f1) oracle pin."""
    import contextlib
    import io
    sd, inputs, keys = train_case(kind, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "v2vnet":
            m = ref_loader.ref_v2vnet_det()
        elif kind == "v2vnet_c2":
            m = ref_loader.ref_v2vnet_det(compress_level=2)
        elif kind == "seg_unet_c3":
            m = ref_loader.ref_seg_unet(compress_level=3)
        elif kind == "fafnet":
            m = ref_loader.ref_fafnet(kd_flag=0)
        elif kind == "when2com":
            m = ref_loader.ref_when2com_det(warp_flag=1)
        elif kind in ("disco", "cat", "agent", "mean", "sum", "max"):
            m = ref_loader.ref_fusion_det(kind)
        elif kind == "seg_when2com":
            m = ref_loader.ref_seg_when2com(num_agent=5, warp_flag=1)
        elif kind == "seg_unet":
            m = ref_loader.ref_seg_unet()
        elif kind in ("seg_mean", "seg_sum", "seg_max", "seg_cat", "seg_agent", "seg_disco"):
            m = ref_loader.ref_fusion_seg(kind[4:], num_agent=5)
        else:
            m = ref_loader.ref_seg_v2vnet(num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.double().train()
    x = inputs[0].double()
    with ref_loader.float64_shim(), ref_loader.cpu_cuda_shim(), contextlib.redirect_stdout(io.StringIO()):
        if kind in ("fafnet", "seg_unet", "seg_unet_c3"):
            r = m(x)
        elif kind == "seg_when2com":
            r = m(x, inputs[1], inputs[2], training=True)
        elif kind == "when2com":
            r = m(x, inputs[1], inputs[2], training=True, MO_flag=True, batch_size=1)
        elif kind in ("seg_v2vnet", "seg_mean", "seg_sum", "seg_max", "seg_cat", "seg_agent", "seg_disco"):
            r = m(x, inputs[1], inputs[2])
        elif kind == "disco":
            r = m(x, inputs[1], inputs[2], batch_size=1)[0]
        elif kind in ("cat", "agent"):
            r = m(x, inputs[1], inputs[2], batch_size=1)
        else:
            r = m(x, inputs[1], inputs[2], batch_size=1)
        if not isinstance(r, dict):
            r = {"logits": r}
        up = make_upstream({k: r[k].shape for k in keys}, seed)
        torch.autograd.backward([r[k] for k in sorted(keys)], [up[k] for k in sorted(keys)])
    out = {"meta": np.asarray([seed], dtype=np.int64), "kind": np.asarray(kind)}
    for name in keys:
        assert r[name].dtype == torch.float64
        out[name + ".shape"] = np.asarray(r[name].shape, dtype=np.int64)
        out[name + ".sub"] = r[name].detach().contiguous().view(-1)[::STRIDE].numpy().copy()
    n_grads = 0
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        n_grads += 1
        out["grad." + k + ".sub"] = p.grad.detach().reshape(-1)[::grad_stride(p.grad.numel())].numpy().copy()
        out["grad." + k + ".norm"] = np.float64(p.grad.detach().norm().item())
    for k, v in m.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            out["bn." + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, tag + ".npz"), **out)
    print(tag, "parameters with gradients:", n_grads)


def gen_options():
    """The when2com constructor options beside the scripts' defaults: has_query=False (ones as queries), sparse=True
    (a no-op in the reference: same fixture values as sparse=False) and det layer=2."""
    gen_when2com("when2com_det_noquery_activated_seed41", 1, 41, 1, "activated", has_query=False)
    gen_when2com("when2com_det_layer2_sparse_activated_seed42_present4", 1, 42, 1, "activated", present=[4], sparse=True,
                 layer=2)
    gen_when2com("when2com_det_layer2_noquery_nowarp_softmax_B2_seed43", 2, 43, 0, "softmax", has_query=False, layer=2)
    gen_seg("seg_when2com_noquery_sparse_activated_seed44", "when2com", 1, 44, inference="activated", warp_flag=1,
            has_query=False, sparse=True)
    # communication at layer 4 (512 ch, 16x16): det When2com (activated: two decoder passes) and det V2VNet (two GNN rounds)
    gen_when2com("when2com_det_layer4_activated_seed48", 1, 48, 1, "activated", layer=4)
    m = ref_loader.ref_v2vnet_det(gnn_iter_times=2, layer=4, layer_channel=512)
    sd = synth.v2vnet_det_state(47, layer_channel=512)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 47, present=[4])
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=1)
    out = {"meta": np.asarray([1, 5, 47, 4], dtype=np.int64)}
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "layer4_v2vnet_det_seed47.npz"), **out)
    print("layer4_v2vnet_det_seed47", out["loc.sum"])
    # parameter-free fusion at layer 4 (the weight nets of Cat / AgentWise / DiscoNet only exist for layers 3 / 2)
    m = ref_loader.ref_fusion_det("sum", layer=4)
    sd = synth.fusion_det_state("sum", 49)
    m.load_state_dict(sd, strict=True)
    m.eval()
    bevs, trans, nat = synth.make_scene(1, 5, 49, present=[3])
    with torch.no_grad():
        r = m(bevs, trans, nat, batch_size=1)
    out = {"meta": np.asarray([1, 5, 49, 4], dtype=np.int64)}
    summarize("loc", r["loc"], out)
    summarize("cls", r["cls"], out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "layer4_sum_det_seed49.npz"), **out)
    print("layer4_sum_det_seed49", out["loc.sum"])


def main():
    if not ref_loader.available():
        print("reference tree not available; golden fixtures can only be generated in the build container")
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "--fusion-only" in sys.argv:
        gen_fusion_all()
        return 0
    if "--compress-only" in sys.argv:
        gen_compress()
        return 0
    if "--layers-only" in sys.argv:
        gen_layers()
        return 0
    if "--options-only" in sys.argv:
        gen_options()
        return 0
    if "--train-only" in sys.argv:
        for tag, (kind, seed) in TRAIN_CASES.items():
            if not os.path.exists(os.path.join(GOLDEN_DIR, tag + ".npz")) or "--force" in sys.argv:
                gen_train_step(tag, kind, seed)
        return 0
    gen_fusion_all()
    gen_compress()
    gen_layers()
    gen_options()
    for tag, (kind, seed) in TRAIN_CASES.items():
        gen_train_step(tag, kind, seed)
    gen_warp("warp_small_seed3", 3)
    gen_convgru("convgru_small_seed4", 4)
    gen_fafnet("fafnet_n2_seed0", 2, 0)
    gen_stages("v2vnet_det_stages_seed0", 0)
    gen_v2vnet("v2vnet_det_A5B1_seed0", 1, 0)
    gen_v2vnet("v2vnet_det_A5B2_seed1_present53", 2, 1, present=[5, 3])
    gen_when2com("when2com_det_warp_activated_seed2", 1, 2, 1, "activated")
    gen_when2com("when2com_det_nowarp_argmax_seed3_present4", 1, 3, 0, "argmax_test", present=[4])
    gen_when2com("when2com_det_warp_softmax_B2_seed4", 2, 4, 1, "softmax", present=[3, 5])
    gen_seg("seg_unet_seed0", "unet", 1, 0)
    gen_seg("seg_v2vnet_seed1_present4", "v2vnet", 1, 1, present=[4])
    gen_seg("seg_when2com_warp_activated_seed2", "when2com", 1, 2, inference="activated", warp_flag=1)
    gen_seg("seg_when2com_nowarp_activated_seed3", "when2com", 1, 3, inference="activated", warp_flag=0, present=[3])  # argmax_test needs torch.cuda.FloatTensor (When2Com_UNet.py:93): not runnable on the CPU reference
    return 0


if __name__ == "__main__":
    sys.exit(main())
