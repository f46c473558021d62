"""Seeded synthetic inputs and weights of the reference's shapes -- the data `bench.py` measures on
and the tests check against (`oracle/synth.py` re-exports these so the golden fixtures and the benchmark use the same
generators).  No reference data or checkpoints exist offline, so:
  * occupancy BEVs   : i.i.d. Bernoulli(p) in {0,1}, agent-major ``[A*B,1,256,256,13]`` float32
                       (shape contract: V2XSimDet.py:291-302, DOC/tutorials/collaborative_models.md:25-50)
  * trans_matrices   : ``[B,A,A,4,4]`` float64, ``T[b,a,k] = P_a^-1 P_k`` for seeded SE(2) poses
                       (nuscenes_pc_util.py:230-232; the model reads ``T[b,j,i]``, DetModelBase.py:158)
  * num_agent_tensor : ``[B,A]`` int64
  * weights          : a seeded state_dict with the reference's key names/shapes (Backbone.py:9-87,
                       DetModelBase.py:268-351, convolutional_rnn/module.py:62-116); He-uniform convs, randomised BN
                       running stats / affine so BN folding is exercised.
Everything comes from CPU generators, so every machine reproduces the same tensors.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

MAP_DIMS = (256, 256, 13)  # Config.py:95-105
NUM_ANCHORS = 6            # Config.py:154-163
CATEGORY_NUM = 2           # Config.py:174-177
BOX_CODE = 6               # Config.py:130-131

# (name, cout, cin, stride) of the 3x3 convs of Backbone (Backbone.py:11-47)
BACKBONE_CONVS = [
    ("conv_pre_1", 32, 13), ("conv_pre_2", 32, 32),
    ("conv1_1", 64, 32), ("conv1_2", 64, 64),
    ("conv2_1", 128, 64), ("conv2_2", 128, 128),
    ("conv3_1", 256, 128), ("conv3_2", 256, 256),
    ("conv4_1", 512, 256), ("conv4_2", 512, 512),
    ("conv5_1", 256, 768), ("conv5_2", 256, 256),
    ("conv6_1", 128, 384), ("conv6_2", 128, 128),
    ("conv7_1", 64, 192), ("conv7_2", 64, 64),
    ("conv8_1", 32, 96), ("conv8_2", 32, 32),
]
BACKBONE_BNS = [
    ("bn_pre_1", 32), ("bn_pre_2", 32),
    ("bn1_1", 64), ("bn1_2", 64), ("bn2_1", 128), ("bn2_2", 128),
    ("bn3_1", 256), ("bn3_2", 256), ("bn4_1", 512), ("bn4_2", 512),
    ("bn5_1", 256), ("bn5_2", 256), ("bn6_1", 128), ("bn6_2", 128),
    ("bn7_1", 64), ("bn7_2", 64), ("bn8_1", 32), ("bn8_2", 32),
]


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)

    def uniform(self, shape, lo, hi):
        return torch.rand(shape, generator=self.g, dtype=torch.float32) * (hi - lo) + lo

    def normal(self, shape, std):
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std


def _conv(sd, g, name, cout, cin, k=(3, 3), gain=6.0):
    fan_in = cin * int(np.prod(k))
    b = math.sqrt(gain / fan_in)
    sd[name + ".weight"] = g.uniform((cout, cin, *k), -b, b)
    sd[name + ".bias"] = g.uniform((cout,), -0.1, 0.1)


def _bn(sd, g, name, c):
    sd[name + ".weight"] = g.uniform((c,), 0.5, 1.5)
    sd[name + ".bias"] = g.normal((c,), 0.1)
    sd[name + ".running_mean"] = g.normal((c,), 0.1)
    sd[name + ".running_var"] = g.uniform((c,), 0.5, 1.5)
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def backbone_state(sd, g, prefix, in_ch=13, compress_level=0):
    """Keys of one full ``Backbone`` (Backbone.py:9-87); both u_encoder and decoder carry all of them."""
    for name, cout, cin in BACKBONE_CONVS:
        if name == "conv_pre_1":
            cin = in_ch
        _conv(sd, g, prefix + name, cout, cin)
    for name, c in BACKBONE_BNS[:2]:
        _bn(sd, g, prefix + name, c)
    for name, c in (("conv3d_1", 64), ("conv3d_2", 128)):
        _conv(sd, g, prefix + name + ".conv3d", c, c, k=(1, 1, 1))
        _bn(sd, g, prefix + name + ".bn3d", c)
    for name, c in BACKBONE_BNS[2:]:
        _bn(sd, g, prefix + name, c)
    if compress_level > 0:
        cc = 256 // (2 ** compress_level)
        _conv(sd, g, prefix + "com_compresser", cc, 256, k=(1, 1))
        _bn(sd, g, prefix + "bn_compress", cc)
        _conv(sd, g, prefix + "com_decompresser", 256, cc, k=(1, 1))
        _bn(sd, g, prefix + "bn_decompress", 256)


def heads_state(sd, g):
    """ClassificationHead / SingleRegressionHead (DetModelBase.py:268-351)."""
    _conv(sd, g, "classification.conv1", 32, 32)
    _conv(sd, g, "classification.conv2", CATEGORY_NUM * NUM_ANCHORS, 32, k=(1, 1), gain=3.0)
    _bn(sd, g, "classification.bn1", 32)
    _conv(sd, g, "regression.box_prediction.0", 32, 32)
    _bn(sd, g, "regression.box_prediction.1", 32)
    _conv(sd, g, "regression.box_prediction.3", NUM_ANCHORS * BOX_CODE, 32, k=(1, 1), gain=3.0)


def v2vnet_det_state(seed=0, layer_channel=256, compress_level=0):
    """state_dict of det V2VNet(config, gnn_iter_times, layer=3, layer_channel=256) (V2VNet.py:14-45)."""
    g = _Gen(seed)
    sd = OrderedDict()
    heads_state(sd, g)
    backbone_state(sd, g, "u_encoder.", compress_level=compress_level)
    backbone_state(sd, g, "decoder.")
    c = layer_channel
    # reference init is U(-1/sqrt(C), 1/sqrt(C)) (module.py:121-124); we keep the gates in
    # their active range with a fan-in scaled draw so sigma/tanh are exercised away from 0.5/0.
    bw = math.sqrt(3.0 / (2 * c * 9))
    sd["convgru.weight_ih_l0"] = g.uniform((3 * c, 2 * c, 3, 3), -bw, bw)
    sd["convgru.weight_hh_l0"] = g.uniform((3 * c, c, 3, 3), -bw, bw)
    sd["convgru.bias_ih_l0"] = g.uniform((3 * c,), -0.5, 0.5)
    sd["convgru.bias_hh_l0"] = g.uniform((3 * c,), -0.5, 0.5)
    return sd


def make_bevs(num_maps, seed=0, p=0.03):
    """``[num_maps,1,256,256,13]`` float32 occupancy in {0,1} (V2XSimDet.py:291-302)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    h, w, z = MAP_DIMS
    return (torch.rand((num_maps, 1, h, w, z), generator=g) < p).to(torch.float32)


def make_poses(batch, num_agent, seed=0, xy=20.0):
    rs = np.random.RandomState(2000 + seed)
    x = rs.uniform(-xy, xy, size=(batch, num_agent))
    y = rs.uniform(-xy, xy, size=(batch, num_agent))
    yaw = rs.uniform(-math.pi, math.pi, size=(batch, num_agent))
    return x, y, yaw


def make_trans_matrices(batch, num_agent, seed=0, present=None):
    """``[B,A,A,4,4]`` float64; ``T[b,a,k] = P_a^-1 P_k`` (nuscenes_pc_util.py:230-232).

    ``present[b]`` = number of non-empty agents of scene b; rows/cols of absent agents are
    zero 4x4s as the dataset pads them (V2XSimDet.py:229-255).
    """
    x, y, yaw = make_poses(batch, num_agent, seed)
    T = np.zeros((batch, num_agent, num_agent, 4, 4), dtype=np.float64)
    for b in range(batch):
        n = num_agent if present is None else int(present[b])
        P = []
        for a in range(num_agent):
            c, s = math.cos(yaw[b, a]), math.sin(yaw[b, a])
            m = np.eye(4)
            m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
            m[0, 3], m[1, 3] = x[b, a], y[b, a]
            P.append(m)
        for a in range(n):
            Pa_inv = np.linalg.inv(P[a])
            for k in range(n):
                T[b, a, k] = Pa_inv @ P[k]
    return torch.from_numpy(T)


def make_scene(batch=1, num_agent=5, seed=0, p=0.03, present=None):
    """Returns (bevs [A*B,1,256,256,13], trans [B,A,A,4,4] f64, num_agent_tensor [B,A] i64).

    Agent-major batch layout: rows ``B*i .. B*(i+1)-1`` = agent i (train_codet.py:287-334).
    """
    bevs = make_bevs(num_agent * batch, seed=seed, p=p)
    nat = torch.full((batch, num_agent), num_agent, dtype=torch.long)
    if present is not None:
        for b in range(batch):
            nat[b, :] = int(present[b])
            for a in range(int(present[b]), num_agent):
                bevs[batch * a + b].zero_()
    trans = make_trans_matrices(batch, num_agent, seed=seed, present=present)
    return bevs, trans, nat


def plant_detections(sd, cls_ref, per_agent=150, loc_scale=0.05, score=0.7):
    """Planted-weight variant of a detection state_dict for the NMS / mAP parity runs.

    With seeded random weights no foreground score passes the reference's hard-coded 0.7 filter
    (postprocess.py:84), so NMS / AP would compare empty sets.  Given the un-planted model's ``cls`` logits
    ``[N, H*W*A, 2]`` for the un-planted state, this returns a copy of ``sd`` whose
      * foreground bias ``classification.conv2.bias[1::2]`` is shifted so that about ``per_agent`` anchors
        per map score above ``score`` (channel = anchor*2 + class, DetModelBase.py:238-245), and
      * last regression layer is scaled by ``loc_scale`` and the bias of every cos-residual channel (code 5 of
        each anchor, channel = anchor*6 + code) is set to 1, so decoded boxes stay near their anchors
        (w = wa / exp(wp), sin/cos = anchor angle rotated by the residual, detection_util.py:385-398) instead
        of spanning the whole map or collapsing to a point.
    Only tensors of the two heads change; backbone / fusion weights are untouched."""
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    d = (cls_ref[..., 1] - cls_ref[..., 0]).reshape(-1).double()
    k = min(d.numel() - 1, per_agent * cls_ref.shape[0])
    q = torch.topk(d, k + 1).values[-1].item()
    shift = math.log(score / (1.0 - score)) - q
    sd["classification.conv2.bias"][1::2] += shift
    sd["regression.box_prediction.3.weight"] *= loc_scale
    sd["regression.box_prediction.3.bias"] *= loc_scale
    sd["regression.box_prediction.3.bias"][5::6] = 1.0
    return sd


# ---- generators the benchmark configs beyond V2VNet det need (FaFNet, seg When2Com_UNet) ------------------------------
def fafnet_state(seed=0, compress_level=0):
    """state_dict of FaFNet (FaFNet.py:17-25; NonIntermediateModelBase.py:24: ``stpn``)."""
    g = _Gen(seed)
    sd = OrderedDict()
    heads_state(sd, g)
    backbone_state(sd, g, "stpn.", compress_level=compress_level)
    return sd


# ---------------------------------------------------------------------------------------------
# segmentation models (CP/models/seg/SegModelBase.py:17-27 channel plan)
# ---------------------------------------------------------------------------------------------
SEG_DOUBLE_CONVS = [  # (prefix, cin, cmid, cout)
    ("inc.double_conv.", 13, 64, 64),
    ("down1.maxpool_conv.1.double_conv.", 64, 128, 128),
    ("down2.maxpool_conv.1.double_conv.", 128, 256, 256),
    ("down3.maxpool_conv.1.double_conv.", 256, 512, 512),
    ("down4.maxpool_conv.1.double_conv.", 512, 512, 512),
    ("up1.conv.double_conv.", 1024, 512, 256),
    ("up2.conv.double_conv.", 512, 256, 128),
    ("up3.conv.double_conv.", 256, 128, 64),
    ("up4.conv.double_conv.", 128, 64, 64),
]


def _double_conv(sd, g, p, cin, cmid, cout):
    _conv(sd, g, p + "0", cmid, cin)
    _bn(sd, g, p + "1", cmid)
    _conv(sd, g, p + "3", cout, cmid)
    _bn(sd, g, p + "4", cout)


def seg_unet_state(seed=0, n_classes=8, g=None, sd=None, compress_level=0):
    """state_dict of seg UNet / the SegModelBase part of every seg model (SegModelBase.py:17-43)."""
    g = g or _Gen(seed)
    sd = OrderedDict() if sd is None else sd
    for p, cin, cmid, cout in SEG_DOUBLE_CONVS:
        _double_conv(sd, g, p, cin, cmid, cout)
    _conv(sd, g, "outc.conv", n_classes, 64, k=(1, 1), gain=3.0)
    if compress_level > 0:
        cc = 512 // (2 ** compress_level)
        _conv(sd, g, "com_compresser", cc, 512, k=(1, 1))
        _bn(sd, g, "bn_compress", cc)
        _conv(sd, g, "com_decompresser", 512, cc, k=(1, 1))
        _bn(sd, g, "bn_decompress", 512)
    return sd


def seg_when2com_state(seed=0, n_classes=8, has_query=True):
    """seg When2Com_UNet (CP/models/seg/When2Com_UNet.py:10-91): SegModelBase + key/query MLPs + attention
    + PolicyNet4 (own inc/down1-3 + conv1..5, :310-339).  ``has_query=False``: the same tensors without ``query_net.*``."""
    g = _Gen(seed)
    sd = seg_unet_state(seed, n_classes, g=g)

    def linear(name, out_f, in_f, gain=3.0):
        b = math.sqrt(gain / in_f)
        sd[name + ".weight"] = g.uniform((out_f, in_f), -b, b)
        sd[name + ".bias"] = g.uniform((out_f,), -0.1, 0.1)

    for net, out in (("key_net", 1024), ("query_net", 32)):
        linear(net + ".fc.0", 256, 4096, gain=6.0)
        linear(net + ".fc.2", 128, 256, gain=6.0)
        linear(net + ".fc.4", out, 128)
    sd["attention_net.linear.weight"] = g.uniform((1024, 32), -0.03, 0.03)
    sd["attention_net.linear.bias"] = g.uniform((1024,), -0.02, 0.02)
    for p, cin, cmid, cout in SEG_DOUBLE_CONVS[:4]:
        _double_conv(sd, g, "query_key_net." + p, cin, cmid, cout)
    for name, cout, cin in (("conv1", 512, 512), ("conv2", 256, 512), ("conv3", 256, 256), ("conv4", 256, 256),
                            ("conv5", 256, 256)):
        _conv(sd, g, "query_key_net.%s.cbr_unit.0" % name, cout, cin)
        _bn(sd, g, "query_key_net.%s.cbr_unit.1" % name, cout)
    if not has_query:
        for k in [k for k in sd if k.startswith("query_net.")]:
            del sd[k]
    if not has_query:
        for k in [k for k in sd if k.startswith("query_net.")]:
            del sd[k]
    return sd


def make_seg_scene(batch=1, num_agent=5, seed=0, p=0.03, present=None):
    """(x [A*B,13,256,256] fp32 NCHW as SegModule.py:49 builds it, trans, num_agent_tensor)."""
    bevs, trans, nat = make_scene(batch, num_agent, seed, p=p, present=present)
    return bevs[:, 0].permute(0, 3, 1, 2).contiguous(), trans, nat


def when2com_det_state(seed=0, has_query=True):
    """state_dict of det When2com(config, layer=3, warp_flag=..) (CP/models/det/When2com.py:22-92):
    V2VNet-style heads / u_encoder / decoder plus key_net / query_net MLPs (:415-430), attention_net.linear
    (:370) and query_key_net = PolicyNet4 (a full Backbone + five Conv2DBatchNormRelu, :335-359).
    ``has_query=False``: the same tensors without ``query_net.*`` (the module is not constructed, :66-69)."""
    g = _Gen(seed)
    sd = OrderedDict()
    heads_state(sd, g)
    backbone_state(sd, g, "u_encoder.")
    backbone_state(sd, g, "decoder.")

    def linear(name, out_f, in_f, gain=3.0):
        b = math.sqrt(gain / in_f)
        sd[name + ".weight"] = g.uniform((out_f, in_f), -b, b)
        sd[name + ".bias"] = g.uniform((out_f,), -0.1, 0.1)

    for net, out in (("key_net", 1024), ("query_net", 32)):
        linear(net + ".fc.0", 256, 4096, gain=6.0)
        linear(net + ".fc.2", 128, 256, gain=6.0)
        linear(net + ".fc.4", out, 128)
    # small so that the 5x5 softmax is neither uniform nor one-hot
    sd["attention_net.linear.weight"] = g.uniform((1024, 32), -0.03, 0.03)
    sd["attention_net.linear.bias"] = g.uniform((1024,), -0.02, 0.02)
    backbone_state(sd, g, "query_key_net.lidar_encoder.")
    for name, cout, cin in (("conv1", 512, 512), ("conv2", 256, 512), ("conv3", 256, 256), ("conv4", 256, 256),
                            ("conv5", 256, 256)):
        _conv(sd, g, "query_key_net.%s.cbr_unit.0" % name, cout, cin)
        _bn(sd, g, "query_key_net.%s.cbr_unit.1" % name, cout)
    if not has_query:
        for k in [k for k in sd if k.startswith("query_net.")]:
            del sd[k]
    return sd
