"""Whole-forward plans of the segmentation models (CP/models/seg/) on the sm_100a kernels.

UNet backbone (SegModelBase.py:17-27): inc / down1..4 (MaxPool2d + DoubleConv) / up1..4 (bilinear x2 with
align_corners=True, cat([skip, up]), DoubleConv) / outc (1x1 -> fp32 NCHW logits).  Every conv + BN(eval) + ReLU is one
v2x_conv_fwd launch; pooling and the bilinear upsample are small byte-mover kernels; the fusion stages reuse the
detection path's warp/mean, ConvGRU, attention and gated-fuse kernels at C = 512.
"""
from __future__ import annotations

import torch

from . import ops
from .nets import DetPlan, _bn
from .ops import EPI_F32_NCHW, EPI_GRU, ConvLaunch

IN_C, IN_C_PAD = 13, 16
H0 = W0 = 256


class DoubleConvW:
    """Packed operands of one DoubleConv (SegModelBase.py:91-106); ``cins`` = channels of the concat sources of conv 0."""

    def __init__(self, sd, p, cins, planes, device):
        self.c0 = ops.pack_conv(sd[p + "0.weight"], sd[p + "0.bias"], _bn(sd, p + "1"), cins=cins, planes=planes,
                                device=device)
        self.c1 = ops.pack_conv(sd[p + "3.weight"], sd[p + "3.bias"], _bn(sd, p + "4"),
                                cins=[sd[p + "3.weight"].shape[1]], planes=planes, device=device)


class SegUNetWeights:
    def __init__(self, sd, planes, device, prefix="", encoder_only=False):
        P = prefix
        self.inc = DoubleConvW(sd, P + "inc.double_conv.", [13], planes, device)
        self.down = [DoubleConvW(sd, P + "down%d.maxpool_conv.1.double_conv." % i,
                                 [sd[P + "down%d.maxpool_conv.1.double_conv.0.weight" % i].shape[1]], planes, device)
                     for i in ((1, 2, 3) if encoder_only else (1, 2, 3, 4))]
        self.compress = None
        if not encoder_only and P + "com_compresser.weight" in sd:   # SegModelBase.py:29-43
            from .nets import pack_compress_pair
            self.compress = pack_compress_pair(sd, P, planes, device)
        if not encoder_only:
            self.up = []
            for i in (1, 2, 3, 4):
                cin = sd[P + "up%d.conv.double_conv.0.weight" % i].shape[1]
                self.up.append(DoubleConvW(sd, P + "up%d.conv.double_conv." % i, [cin // 2, cin // 2], planes, device))
            n_cls = sd[P + "outc.conv.weight"].shape[0]
            self.outc = ops.pack_conv(sd[P + "outc.conv.weight"], sd[P + "outc.conv.bias"], None,
                                      cins=[sd[P + "outc.conv.weight"].shape[1]], planes=planes, device=device,
                                      cout_pad=((n_cls + 31) // 32) * 32)


class SegPlan(DetPlan):
    """UNet machinery on top of the shared plan / CUDA-graph plumbing."""

    def build_input(self):
        self.x_in_f32 = torch.zeros((self.n, IN_C, H0, W0), dtype=torch.float32, device=self.device)
        x_in = self.act("x_in", H0, W0, IN_C_PAD)
        src, planes = self.x_in_f32, self.planes
        self.add(lambda: ops.pack_input_nchw(src, IN_C_PAD, planes, out=x_in))
        return x_in

    def double_conv(self, w: DoubleConvW, srcs, name):
        t = self.conv(w.c0, srcs, name + "a")
        return self.conv(w.c1, [t], name)

    def pool(self, x, name):
        out = self.act(name, x.shape[2] // 2, x.shape[3] // 2, x.shape[4])
        self.add(lambda: ops.maxpool2(x, out=out))
        return out

    def upsample(self, x, name):
        out = self.act(name, x.shape[2] * 2, x.shape[3] * 2, x.shape[4])
        self.add(lambda: ops.upsample_bilinear2(x, out=out))
        return out

    def build_encoder(self, w: SegUNetWeights, x_in, tag=""):
        x1 = self.double_conv(w.inc, [x_in], tag + "x1")
        x2 = self.double_conv(w.down[0], [self.pool(x1, tag + "p1")], tag + "x2")
        x3 = self.double_conv(w.down[1], [self.pool(x2, tag + "p2")], tag + "x3")
        x4 = self.double_conv(w.down[2], [self.pool(x3, tag + "p3")], tag + "x4")
        if w.compress is not None:   # UNet.py:30-32, seg/V2VNet.py:32-34, When2Com_UNet.py:163-165, seg/FusionBase.py:31-33
            t = self.conv(w.compress[0], [x4], tag + "x4c")
            x4 = self.conv(w.compress[1], [t], tag + "x4d")
        return x1, x2, x3, x4

    def build_decoder(self, w: SegUNetWeights, feat, x1, x2, x3):
        """down4 / up1..4 / outc on the (fused) layer-4 map; cat order is [skip, upsampled] (SegModelBase.py:141)."""
        x5 = self.double_conv(w.down[3], [self.pool(feat, "p4")], "x5")
        t = self.double_conv(w.up[0], [feat, self.upsample(x5, "u1")], "x6")
        t = self.double_conv(w.up[1], [x3, self.upsample(t, "u2")], "x7")
        t = self.double_conv(w.up[2], [x2, self.upsample(t, "u3")], "x8")
        t = self.double_conv(w.up[3], [x1, self.upsample(t, "u4")], "x9")
        n_cls = w.outc.cout
        self.logits = torch.empty((self.n, n_cls, H0, W0), dtype=torch.float32, device=self.device)
        self.add(ConvLaunch(w.outc, [t], epilogue=EPI_F32_NCHW, relu=False, out0=self.logits,
                            block_n=32))

    def set_x(self, x):
        self.x_in_f32.copy_(x.reshape(self.x_in_f32.shape), non_blocking=True)

    def result(self):
        return self.logits


class SegUNetPlan(SegPlan):
    """seg UNet.forward (UNet.py:24-44): no fusion."""

    def __init__(self, sd, n_maps, planes=1, device="cuda"):
        super().__init__(n_maps, planes, device)
        planes = self.planes   # ``planes`` may have been a mode name / Precision (v2x_b200/precision.py)
        ops.require_gpu()
        self.w = SegUNetWeights(sd, planes, self.device)
        x_in = self.build_input()
        x1, x2, x3, x4 = self.build_encoder(self.w, x_in)
        self.x4 = x4
        self.build_decoder(self.w, x4, x1, x2, x3)

    def forward(self, x):
        self.set_x(x)
        self.run()
        return self.result()


class _FusedSegPlan(SegPlan):
    def _common(self, batch, agents):
        dev = self.device
        self.batch, self.agents = batch, agents
        self.trans = torch.zeros((batch, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch, agents), agents, dtype=torch.int64, device=dev)

    def forward(self, x, trans_matrices, num_agent_tensor):
        self.set_x(x)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)
        self.run()
        return self.result()


class SegV2VNetPlan(_FusedSegPlan):
    """seg V2VNet.forward (seg/V2VNet.py:25-92): one GNN round at C = 512, neighbour mean includes self."""

    def __init__(self, sd, batch, agents=5, planes=1, device="cuda", only_v2i=False):
        super().__init__(batch * agents, planes, device)
        planes = self.planes   # ``planes`` may have been a mode name / Precision (v2x_b200/precision.py)
        ops.require_gpu()
        self._common(batch, agents)
        self.w = SegUNetWeights(sd, planes, self.device)
        self.gru_h, self.gru_m = ops.pack_gru_split(sd["convgru.weight_ih_l0"], sd["convgru.bias_ih_l0"],
                                                    sd["convgru.bias_hh_l0"], planes=planes, device=self.device,
                                                    mmas=self.prec.mmas("gru"))
        x_in = self.build_input()
        x1, x2, x3, x4 = self.build_encoder(self.w, x_in)
        c4 = x4.shape[-1]
        mean = self.act("mean", 32, 32, c4)
        trans, na = self.trans, self.num_agent
        self.add(lambda: ops.warp_mean(x4, trans, na, batch, agents, include_self=True, only_v2i=only_v2i, out=mean))
        fused = self.build_gru_rounds(x4, mean, 1, batch, agents, 0)
        self.build_decoder(self.w, fused, x1, x2, x3)


class SegWhen2comPlan(_FusedSegPlan):
    """seg When2Com_UNet.forward (When2Com_UNet.py:144-307), incl. the key/query row quirk.
    ``has_query=False``: every agent's query is a vector of ones (:219-225); the query MLP is not run."""

    def __init__(self, sd, batch, agents=5, planes=1, device="cuda", warp_flag=1, inference="activated",
                 training=False, only_v2i=False, has_query=True):
        super().__init__(batch * agents, planes, device)
        planes = self.planes   # ``planes`` may have been a mode name / Precision (v2x_b200/precision.py)
        ops.require_gpu()
        dev = self.device
        self._common(batch, agents)
        f32 = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.w = SegUNetWeights(sd, planes, dev)
        self.pol_w = SegUNetWeights(sd, planes, dev, prefix="query_key_net.", encoder_only=True)
        self.pol_convs = []
        for name, stride, cin in (("conv1", 1, 512), ("conv2", 1, 512), ("conv3", 2, 256), ("conv4", 1, 256),
                                  ("conv5", 2, 256)):
            pre = "query_key_net.%s.cbr_unit." % name
            self.pol_convs.append(ops.pack_conv(sd[pre + "0.weight"], sd[pre + "0.bias"], _bn(sd, pre + "1"), cins=[cin],
                                                stride=stride, planes=planes, device=dev))
        self.mlp = {net: [(f32("%s.fc.%d.weight" % (net, i)), f32("%s.fc.%d.bias" % (net, i))) for i in (0, 2, 4)]
                    for net in (("key_net", "query_net") if has_query else ("key_net",))}
        self.att_w, self.att_b = f32("attention_net.linear.weight"), f32("attention_net.linear.bias")
        trans, na, n = self.trans, self.num_agent, self.n

        x_in = self.build_input()
        x1, x2, x3, x4 = self.build_encoder(self.w, x_in)
        t = self.build_encoder(self.pol_w, x_in, tag="pol_")[3]
        for i, pc in enumerate(self.pol_convs):
            t = self.conv(pc, [t], "pol_c%d" % (i + 1))
        qk = t   # [N, 8, 8, 256]: 16384 features per map, viewed as 4 rows of 4096 (Q9); rows 0..N-1 are used
        feats = {"query_net": torch.ones((n, self.att_w.shape[1]), dtype=torch.float32, device=dev)}
        for net in self.mlp:
            (w0, b0), (w1, b1), (w2, b2) = self.mlp[net]
            h0 = torch.empty((n, w0.shape[0]), dtype=torch.float32, device=dev)
            h1 = torch.empty((n, w1.shape[0]), dtype=torch.float32, device=dev)
            h2 = torch.empty((n, w2.shape[0]), dtype=torch.float32, device=dev)
            self.add(lambda w0=w0, b0=b0, h0=h0: ops.linear(qk, w0, b0, relu=True, out=h0, act_input=True, rows=n))
            self.add(lambda w1=w1, b1=b1, h0=h0, h1=h1: ops.linear(h0, w1, b1, relu=True, out=h1))
            self.add(lambda w2=w2, b2=b2, h1=h1, h2=h2: ops.linear(h1, w2, b2, relu=False, out=h2))
            feats[net] = h2
        self.attn = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
        self.coef = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
        gate = 0 if (training or inference == "softmax") else ops.GATE_MODES[inference]
        keys, querys, attn, coef, aw, ab = feats["key_net"], feats["query_net"], self.attn, self.coef, self.att_w, self.att_b
        self.add(lambda: ops.attn_scores(keys, querys, aw, ab, batch, agents, gate, attn=attn, coef=coef))
        fused = self.act("fused", 32, 32, x4.shape[-1])
        self.add(lambda: ops.warp_gated(x4, trans, na, coef, batch, agents, warp_flag=warp_flag, only_v2i=only_v2i,
                                        out=fused))
        self.build_decoder(self.w, fused, x1, x2, x3)


class SegFusionPlan(_FusedSegPlan):
    """seg intermediate-fusion baselines (seg/FusionBase.py:25-84): UNet encoder -> FuseStage(kind) on the 512-channel
    layer-4 map -> down4 / up1..4 / outc."""

    def __init__(self, sd, kind, batch, agents=5, planes=1, device="cuda", only_v2i=False):
        from .nets import FuseStage
        super().__init__(batch * agents, planes, device)
        planes = self.planes   # ``planes`` may have been a mode name / Precision (v2x_b200/precision.py)
        ops.require_gpu()
        self._common(batch, agents)
        self.kind = kind
        self.w = SegUNetWeights(sd, planes, self.device)
        x_in = self.build_input()
        x1, x2, x3, x4 = self.build_encoder(self.w, x_in)
        self.fuse = FuseStage(self, kind, sd, x4, self.trans, self.num_agent, batch, agents, only_v2i=only_v2i, seg=True)
        self.fused = self.fuse.out
        self.build_decoder(self.w, self.fused, x1, x2, x3)

    def kd_layers(self):
        """(x9, x8, x7, x6, x5, feat_mat) as fp32 NCHW (seg/FusionBase.py:81-82)."""
        f = ops.act_to_float
        return tuple(f(self.ws[k]) for k in ("x9", "x8", "x7", "x6", "x5")) + (f(self.fused),)
