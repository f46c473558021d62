"""Whole-forward plans of the detection models on the sm_100a kernels.

A plan owns the packed operands (derived from a reference-format ``state_dict``), a static
workspace of act tensors for a fixed (batch, agents) geometry, and the bound launch list of the
forward.  ``run()`` replays the launches on the current stream; ``capture()`` records them into
one CUDA graph (no python, no host syncs, no per-call H2D in the replayed step -- the reference's
python triple loop / per-warp ``torch.tensor(mask)`` / ``.item()`` syncs, V2VNet.py:66-107 and
DetModelBase.py:163, disappear).

Layer order and semantics follow CP/models/det/backbone/Backbone.py:89-242,
CP/models/det/V2VNet.py:47-120 and CP/models/det/base/DetModelBase.py:226-265.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops, precision
from .ops import EPI_ACT, EPI_F32_SPLIT, EPI_GRU, EPI_TAIL_F32_SPLIT, ConvLaunch, PackedConv

import os

H0 = W0 = 256
# the two head convs run as one launch (EPI_TAIL_F32_SPLIT); V2X_NO_FUSE_HEADS=1 keeps the two-launch form (A/B tests)
FUSE_HEADS = not os.environ.get("V2X_NO_FUSE_HEADS")
# ConvGRU pre-activations ride the rounds' GEMM as a bf16 second source (gru_pre_act); V2X_GRU_ADD_F32=1 keeps the fp32
# epilogue-add form (A/B tests)
GRU_PRE_ACT = not os.environ.get("V2X_GRU_ADD_F32")
IN_C, IN_C_PAD = 13, 16


def _bn(sd, name):
    return (sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"], sd[name + ".running_var"])


def pack_compress_pair(sd, prefix, planes, device):
    """com_compresser (C -> C / 2^level, 1x1 + BN + ReLU) and com_decompresser (back to C) as two 1x1 conv operands.
    The compressed width can be as small as 1 channel; it is zero-padded to 32 (zero weight rows with identity BN give
    relu(0) = 0, and the decompresser's matching input columns are zero), which leaves the arithmetic unchanged."""
    wc, bc = sd[prefix + "com_compresser.weight"], sd[prefix + "com_compresser.bias"]
    wd, bd = sd[prefix + "com_decompresser.weight"], sd[prefix + "com_decompresser.bias"]
    cc, c = wc.shape[0], wc.shape[1]
    ccp = max(32, cc)
    g, b, m, v = _bn(sd, prefix + "bn_compress")
    if ccp != cc:
        z = lambda t, fill: torch.cat([t, torch.full((ccp - cc,), fill, dtype=t.dtype, device=t.device)])  # noqa: E731
        wc = torch.cat([wc, wc.new_zeros((ccp - cc,) + tuple(wc.shape[1:]))], 0)
        bc, g, b, m, v = z(bc, 0.0), z(g, 1.0), z(b, 0.0), z(m, 0.0), z(v, 1.0)
        wd = torch.cat([wd, wd.new_zeros((wd.shape[0], ccp - cc) + tuple(wd.shape[2:]))], 1)
    comp = ops.pack_conv(wc, bc, (g, b, m, v), cins=[c], planes=planes, device=device)
    decomp = ops.pack_conv(wd, bd, _bn(sd, prefix + "bn_decompress"), cins=[ccp], planes=planes, device=device)
    return comp, decomp


class BackboneWeights:
    """Packed operands of the used halves of a reference ``Backbone`` (encoder and/or decoder)."""

    ENC = [("conv_pre_1", "bn_pre_1", 1, [13]), ("conv_pre_2", "bn_pre_2", 1, [32]),
           ("conv1_1", "bn1_1", 2, [32]), ("conv1_2", "bn1_2", 1, [64]),
           ("conv2_1", "bn2_1", 2, [64]), ("conv2_2", "bn2_2", 1, [128]),
           ("conv3_1", "bn3_1", 2, [128]), ("conv3_2", "bn3_2", 1, [256]),
           ("conv4_1", "bn4_1", 2, [256]), ("conv4_2", "bn4_2", 1, [512])]
    DEC = [("conv5_1", "bn5_1", 1, [512, 256]), ("conv5_2", "bn5_2", 1, [256]),
           ("conv6_1", "bn6_1", 1, [256, 128]), ("conv6_2", "bn6_2", 1, [128]),
           ("conv7_1", "bn7_1", 1, [128, 64]), ("conv7_2", "bn7_2", 1, [64]),
           ("conv8_1", "bn8_1", 1, [64, 32]), ("conv8_2", "bn8_2", 1, [32])]

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, planes: int, device, encoder=True, decoder=True):
        self.c: Dict[str, PackedConv] = {}
        prec = precision.resolve(planes)     # an int (1 / 2), a mode name or a Precision: per-layer MMA passes
        planes = prec.planes
        layers = (self.ENC if encoder else []) + (self.DEC if decoder else [])
        for conv, bn, stride, cins in layers:
            self.c[conv] = ops.pack_conv(sd[prefix + conv + ".weight"], sd[prefix + conv + ".bias"],
                                         _bn(sd, prefix + bn), cins=cins, stride=stride, planes=planes, device=device,
                                         mmas=prec.mmas(conv))
        if encoder and prefix + "com_compresser.weight" in sd:
            # optional compress / decompress of the communicated layer x_3 (Backbone.py:74-87,138-141)
            self.c["com_compresser"], self.c["com_decompresser"] = pack_compress_pair(sd, prefix, planes, device)
        if encoder:
            for name in ("conv3d_1", "conv3d_2"):
                self.c[name] = ops.pack_conv(sd[prefix + name + ".conv3d.weight"], sd[prefix + name + ".conv3d.bias"],
                                             _bn(sd, prefix + name + ".bn3d"), cins=[sd[prefix + name + ".conv3d.weight"].shape[1]],
                                             planes=planes, device=device, mmas=prec.mmas(name))


class HeadWeights:
    def __init__(self, sd, planes, device):
        prec = precision.resolve(planes)
        planes = prec.planes
        self.head1, self.head2, self.n_cls = ops.pack_heads(
            sd["classification.conv1.weight"], sd["classification.conv1.bias"], _bn(sd, "classification.bn1"),
            sd["regression.box_prediction.0.weight"], sd["regression.box_prediction.0.bias"],
            _bn(sd, "regression.box_prediction.1"),
            sd["classification.conv2.weight"], sd["classification.conv2.bias"],
            sd["regression.box_prediction.3.weight"], sd["regression.box_prediction.3.bias"],
            planes=planes, device=device, mmas=prec.mmas("heads"))


class DetPlan:
    """Shared machinery: workspace, launch list, graph capture."""

    def __init__(self, n_maps: int, planes: int, device):
        self.prec = precision.resolve(planes)
        self.n, self.planes, self.device = n_maps, self.prec.planes, torch.device(device)
        self.launches: List = []
        self.ws: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.flops = 0.0
        self.n_kernels = 0
        # optional fork/join: launches [side_lo, side_hi) run on a second stream, concurrently with the launches after
        # them on the main stream, and are joined before launch `side_join` (see V2VNetDetPlan: the x_4 encoder branch
        # runs beside the warp / first GRU launch, which leave most of the SMs' tensor pipes idle)
        self.side_lo = self.side_hi = self.side_join = -1
        self._side_stream = None

    def act(self, name, h, w, c):
        t = ops.empty_act(self.planes, self.n, h, w, c, self.device)
        self.ws[name] = t
        return t

    def add(self, launch):
        self.launches.append(launch)
        self.n_kernels += 1
        self.flops += getattr(launch, "flops", 0.0)

    def conv(self, pc, srcs, out_name, *, upsample2x=False, **kw):
        planes, n, h_in, w_in, _ = srcs[0].shape
        up = 2 if upsample2x else 1
        out = self.act(out_name, h_in // pc.stride * up, w_in // pc.stride * up, pc.cout)
        self.add(ConvLaunch(pc, srcs, out0=out, upsample2x=upsample2x, **kw))
        return out

    def upsample2(self, x, out_name):
        """Nearest 2x upsampling of an act through the conv kernel's upsampling store: a 1x1 identity operand without
        ReLU.  Exact in every mode (bf16: v * 1; fp16 hi/lo: hi * 1 + lo * 1 in the fp32 accumulator, re-split on store).
        Used where a FUSED layer-4 map feeds the decoder, which reads layer 4 only through F.interpolate(x_4, 2)
        (Backbone.py:176); the un-fused x_4 is stored upsampled by its producer instead (build_encoder)."""
        c = x.shape[-1]
        eye = torch.eye(c, dtype=torch.float32, device=self.device).view(c, c, 1, 1)
        pc = ops.pack_conv(eye, None, None, cins=[c], planes=self.planes, device=self.device,
                           mmas=3 if self.planes == 2 else 1)
        return self.conv(pc, [x], out_name, relu=False, upsample2x=True)

    def check_voxels(self):
        """Voxel mode: raise (like numpy's IndexError at V2XSimDet.py:296) if any uploaded row was out of range.
        Reads one device int32 -- a host sync, so callers on a latency-critical loop may skip it."""
        bad = int(self.vox_bad.item())
        if bad:
            self.vox_bad.zero_()
            raise IndexError("%d voxel rows were outside the %dx%dx%d grid / map range" % (bad, H0, W0, IN_C))

    # ---- encoder / decoder / heads (Backbone.py:89-242, DetModelBase.py:226-265) ----
    def build_input(self, mode="f32", voxel_capacity=0):
        """The first launch of every plan: the caller's input -> act [P, N, 256, 256, 16].
        mode "f32": fp32 BEV [N,1,256,256,13] (what the reference scripts hand the model, V2VNet.py:47-51);
        mode "u8":  the same grid as bool/uint8 (V2XSimDet.py:299 before .astype(np.float32)): 4x fewer H2D bytes;
        mode "voxels": sparse voxel rows (map, i0, i1, i2) int32, scattered + rot90'd on device (V2XSimDet.py:294-299,
        SURVEY 8(f3)); ``voxel_capacity`` rows are reserved, the live count is a device scalar."""
        self.input_mode = mode
        x_in = self.act("x_in", H0, W0, IN_C_PAD)
        planes = self.planes
        if mode == "f32":
            self.bev_in = torch.zeros((self.n, 1, H0, W0, IN_C), dtype=torch.float32, device=self.device)
            bev = self.bev_in
            self.add(lambda: ops.pack_input(bev, IN_C_PAD, planes, out=x_in))
        elif mode == "u8":
            self.bev_in = torch.zeros((self.n, 1, H0, W0, IN_C), dtype=torch.uint8, device=self.device)
            bev = self.bev_in
            self.add(lambda: ops.pack_input_u8(bev, IN_C_PAD, planes, out=x_in))
        elif mode == "voxels":
            cap = int(voxel_capacity) or 32768 * self.n
            self.vox_idx = torch.zeros((cap, 4), dtype=torch.int32, device=self.device)
            self.vox_count = torch.zeros((1,), dtype=torch.int32, device=self.device)
            self.vox_bad = torch.zeros((1,), dtype=torch.int32, device=self.device)
            idx, cnt, bad = self.vox_idx, self.vox_count, self.vox_bad
            self.add(lambda: ops.voxelize(idx, cnt, x_in, IN_C, bad, rot90=True))
        else:
            raise ValueError("input mode must be f32, u8 or voxels")
        return x_in

    def set_bevs(self, bevs):
        """Async copy of the step's input into the plan's static buffer: a dense BEV (modes f32 / u8) or, in voxel
        mode, an int32 [n, 4] tensor of (map, i0, i1, i2) rows."""
        if self.input_mode == "voxels":
            n = int(bevs.shape[0])
            if n > self.vox_idx.shape[0]:
                raise ops.V2XError("%d voxel rows exceed the plan's capacity %d" % (n, self.vox_idx.shape[0]))
            assert bevs.dtype == torch.int32 and bevs.dim() == 2 and bevs.shape[1] == 4
            self.vox_idx[:n].copy_(bevs, non_blocking=True)
            self.vox_count.fill_(n)
        else:
            self.bev_in.copy_(bevs.reshape(self.bev_in.shape), non_blocking=True)

    def build_encoder(self, w: BackboneWeights, x_in, tag="", upsample_x4=True):
        c = w.c
        t = self.conv(c["conv_pre_1"], [x_in], tag + "x0a")
        x0 = self.conv(c["conv_pre_2"], [t], tag + "x0")
        t = self.conv(c["conv1_1"], [x0], tag + "x1a")
        t = self.conv(c["conv1_2"], [t], tag + "x1b")
        x1 = self.conv(c["conv3d_1"], [t], tag + "x1")
        t = self.conv(c["conv2_1"], [x1], tag + "x2a")
        t = self.conv(c["conv2_2"], [t], tag + "x2b")
        x2 = self.conv(c["conv3d_2"], [t], tag + "x2")
        t = self.conv(c["conv3_1"], [x2], tag + "x3a")
        x3 = self.conv(c["conv3_2"], [t], tag + "x3")
        self.x4_branch = (len(self.launches), len(self.launches) + 2)   # launch indices of conv4_1, conv4_2
        t = self.conv(c["conv4_1"], [x3], tag + "x4a")
        # in the detection decoder x_4 is only ever consumed through F.interpolate(x_4, 2) (Backbone.py:176):
        # store it upsampled.  (PolicyNet4 consumes the plain x_4, When2com.py:354.)
        x4 = self.conv(c["conv4_2"], [t], tag + ("x4u" if upsample_x4 else "x4"), upsample2x=upsample_x4)
        if "com_compresser" in c:   # x_4 was computed from the uncompressed x_3 (Backbone.py:131-141)
            t = self.conv(c["com_compresser"], [x3], tag + "x3c")
            x3 = self.conv(c["com_decompresser"], [t], tag + "x3d")
        return x0, x1, x2, x3, x4

    def build_decoder(self, w: BackboneWeights, x0, x1, x2, x3, x4u, tag=""):
        c = w.c
        t = self.conv(c["conv5_1"], [x4u, x3], tag + "x5a")
        x5u = self.conv(c["conv5_2"], [t], tag + "x5u", upsample2x=True)
        t = self.conv(c["conv6_1"], [x5u, x2], tag + "x6a")
        x6u = self.conv(c["conv6_2"], [t], tag + "x6u", upsample2x=True)
        t = self.conv(c["conv7_1"], [x6u, x1], tag + "x7a")
        x7u = self.conv(c["conv7_2"], [t], tag + "x7u", upsample2x=True)
        t = self.conv(c["conv8_1"], [x7u, x0], tag + "x8a")
        return self.conv(c["conv8_2"], [t], tag + "x8")

    def build_gru_rounds(self, x3, mean, gnn_iter, batch, agents, map_offset):
        """``gnn_iter`` zero-hidden ConvGRU rounds on cat([h, mean]) (V2VNet.py:99-101).  The mean half of the input never
        changes between rounds (neighbours are always warped from the original maps), so its contribution
        conv(mean, W_ih[:, C:]) + bias is computed once into fp32 pre-activations and added in each round's epilogue."""
        c3, hh, ww = x3.shape[-1], x3.shape[2], x3.shape[3]
        if self.gru_h.gru_pre_act:
            # pre-activations as a bf16 act tensor, accumulated by the rounds' own GEMM through identity weight columns
            self.gru_pre = self.act("gru_pre", hh, ww, 3 * c3)
            self.add(ConvLaunch(self.gru_m, [mean], epilogue=EPI_ACT, relu=False, out0=self.gru_pre))
        else:
            self.gru_pre = torch.empty((self.n, hh, ww, 3 * c3), dtype=torch.float32, device=self.device)
            self.add(ConvLaunch(self.gru_m, [mean], epilogue=EPI_F32_SPLIT, relu=False, out0=self.gru_pre, split=3 * c3))
        h = x3
        for r in range(gnn_iter):
            out = self.act("h%d" % (r + 1), hh, ww, c3)
            if self.gru_h.gru_pre_act:
                self.add(ConvLaunch(self.gru_h, [h, self.gru_pre], epilogue=EPI_GRU, out0=out, passthrough=x3,
                                    num_agent=self.num_agent, batch=batch, agents=agents, map_offset=map_offset))
            else:
                self.add(ConvLaunch(self.gru_h, [h], epilogue=EPI_GRU, out0=out, passthrough=x3, num_agent=self.num_agent,
                                    batch=batch, agents=agents, map_offset=map_offset, gru_add=self.gru_pre))
            h = out
        return h

    def build_heads(self, hw: HeadWeights, x8):
        """cls / reg heads (DetModelBase.py:283-296, 319-329): conv3x3+BN+ReLU (both heads stacked, 64 rows) and the
        block-diagonal 1x1 as ONE launch -- the 64-channel intermediate lives only in shared memory."""
        n_cls = hw.n_cls
        n_loc = hw.head2.cout - n_cls
        self.cls = torch.empty((self.n, H0, W0, n_cls), dtype=torch.float32, device=self.device)
        self.loc = torch.empty((self.n, H0, W0, n_loc), dtype=torch.float32, device=self.device)
        if FUSE_HEADS and hw.head1.cout == 64 and hw.head2.cout_pad <= 64 and hw.head2.cout_pad % 16 == 0:
            self.add(ConvLaunch(hw.head1, [x8], epilogue=EPI_TAIL_F32_SPLIT, relu=True, out0=self.cls, out1=self.loc,
                                split=n_cls, block_n=64, tail=hw.head2))
            return
        t = self.conv(hw.head1, [x8], "head1")
        self.add(ConvLaunch(hw.head2, [t], epilogue=EPI_F32_SPLIT, relu=False, out0=self.cls, out1=self.loc,
                            split=n_cls, block_n=hw.head2.cout))

    # ---- execution ----
    def _issue(self):
        """Issue the launch list on the current stream, forking [side_lo, side_hi) onto the side stream."""
        if self.side_lo < 0 or os.environ.get("V2X_NO_SIDE_STREAM"):
            for l in self.launches:
                l()
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        main, side = torch.cuda.current_stream(), self._side_stream
        for i, l in enumerate(self.launches):
            if i == self.side_lo:
                side.wait_stream(main)          # fork: the side branch sees everything issued so far
            if i == self.side_join:
                main.wait_stream(side)          # join
            if self.side_lo <= i < self.side_hi:
                with torch.cuda.stream(side):
                    l()
            else:
                l()
        if self.side_join >= len(self.launches) or self.side_join < 0:
            main.wait_stream(side)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._issue()

    def capture(self):
        """Record the whole forward into one CUDA graph (side stream warm-up first, as CUDA requires)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._issue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._issue()
        self.graph = g

    def result(self):
        """Reference output contract (DetModelBase.py:238-252): cls [N, H*W*6, 2], loc [N,H,W,6,1,6]."""
        n = self.n
        return {"loc": self.loc.view(n, H0, W0, 6, 1, 6), "cls": self.cls.view(n, -1, 2)}


class V2VNetDetPlan(DetPlan):
    """det V2VNet forward (V2VNet.py:47-120) for ``batch`` scenes x ``agents`` agent slots."""

    def __init__(self, sd, batch: int, agents: int = 5, gnn_iter: int = 3, planes: int = 1, device="cuda",
                 only_v2i=False, input_mode="f32", voxel_capacity=0, layer: int = 3):
        super().__init__(batch * agents, planes, device)
        prec, planes = self.prec, self.prec.planes
        if layer not in (1, 2, 3, 4):
            raise ops.V2XError("V2VNet on the sm_100a path communicates at layer 1..4 (the ConvGRU tile needs a multiple "
                               "of 64 channels: layer 0 has 32)")
        ops.require_gpu()
        self.batch, self.agents, self.gnn_iter = batch, agents, gnn_iter
        dev = self.device
        self.enc_w = BackboneWeights(sd, "u_encoder.", prec, dev, encoder=True, decoder=False)
        self.dec_w = BackboneWeights(sd, "decoder.", prec, dev, encoder=False, decoder=True)
        self.head_w = HeadWeights(sd, prec, dev)
        # (the 512-channel layer-4 GRU takes the fp32 pre-activation form, like the seg V2VNet's 512-channel GRU)
        self.gru_h, self.gru_m = ops.pack_gru_split(sd["convgru.weight_ih_l0"], sd["convgru.bias_ih_l0"],
                                                    sd["convgru.bias_hh_l0"], planes=planes, device=dev,
                                                    pre_act=GRU_PRE_ACT and planes == 1 and layer != 4,
                                                    mmas=prec.mmas("gru"))
        self.trans = torch.zeros((batch, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch, agents), agents, dtype=torch.int64, device=dev)

        x_in = self.build_input(input_mode, voxel_capacity)
        # layer 4: the GNN block needs the plain 16x16 x_4; the decoder gets the fused map upsampled afterwards
        x0, x1, x2, x3, x4u = self.build_encoder(self.enc_w, x_in, upsample_x4=(layer != 4))
        xs = [x0, x1, x2, x3, x4u]
        xl = xs[layer]              # the communicated layer (DetModelBase.get_feature_maps_and_size, :71-92)
        cl, hl, wl = xl.shape[-1], xl.shape[2], xl.shape[3]
        # neighbours are always warped from the ORIGINAL encoder maps (V2VNet.py:85-94), so the mean is
        # round-invariant: one warp launch per frame instead of 60 grid_sample calls
        mean = self.act("mean", hl, wl, cl)
        trans, na = self.trans, self.num_agent
        self.add(lambda: ops.warp_mean(xl, trans, na, batch, agents, include_self=False, only_v2i=only_v2i, out=mean))
        xs[layer] = self.build_gru_rounds(xl, mean, gnn_iter, batch, agents, 0)
        # conv4_1 / conv4_2 only feed the decoder: they run on the side stream beside the warp + GRU launches and
        # are joined before conv5_1 (layer 3 without compression; otherwise the fuse input is produced after them)
        if layer == 3 and "com_compresser" not in self.enc_w.c:
            self.side_lo, self.side_hi = self.x4_branch
            self.side_join = len(self.launches)
        if layer == 4:
            xs[4] = self.upsample2(xs[4], "x4u")
        x8 = self.build_decoder(self.dec_w, xs[0], xs[1], xs[2], xs[3], xs[4])
        self.build_heads(self.head_w, x8)

    def set_inputs(self, bevs, trans_matrices, num_agent_tensor):
        """Async copies into the plan's static input buffers (host or device sources)."""
        self.set_bevs(bevs)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)

    def forward(self, bevs, trans_matrices, num_agent_tensor):
        self.set_inputs(bevs, trans_matrices, num_agent_tensor)
        self.run()
        return self.result()


class FaFNetPlan(DetPlan):
    """FaFNet / STPN forward: encoder -> decoder -> heads, no fusion (FaFNet.py:28-39)."""

    def __init__(self, sd, n_maps: int, planes: int = 1, device="cuda", heads: bool = True, input_mode="f32",
                 voxel_capacity=0):
        super().__init__(n_maps, planes, device)
        prec, planes = self.prec, self.prec.planes
        ops.require_gpu()
        self.w = BackboneWeights(sd, "stpn.", prec, self.device, encoder=True, decoder=True)
        x_in = self.build_input(input_mode, voxel_capacity)
        x0, x1, x2, x3, x4u = self.build_encoder(self.w, x_in)
        x8 = self.build_decoder(self.w, x0, x1, x2, x3, x4u)
        self.has_heads = heads   # TeacherNet.forward returns the STPN layers only (TeacherNet.py:10-13)
        if heads:
            self.head_w = HeadWeights(sd, prec, self.device)
            self.build_heads(self.head_w, x8)

    def forward(self, bevs):
        self.set_bevs(bevs)
        self.run()
        return self.result() if self.has_heads else None


class When2comDetPlan(DetPlan):
    """det When2com / who2com forward in eval mode (CP/models/det/When2com.py:150-332, MO_flag=True).

    encoder -> [policy encoder + 5 convs -> key/query MLPs -> attention scores] -> gated fuse -> decoder -> (eval:
    re-gated fuse -> second decoder pass whose layer-0 skip is the first pass's output) -> heads.

    ``has_query=False``: every agent's query is a vector of ones (When2com.py:241-245) -- the query MLP is not run and
    the attention kernel reads a constant buffer.  ``layer`` 2 or 3: the communicated encoder layer (:167-190); the
    reference's argmax_test branch is written for layer 3 only (:289-291 hands the fused map to the decoder's layer-3
    slot and crashes on a layer-2 map, profiles/r02_reference_option_probe.txt), so that combination is refused."""

    def __init__(self, sd, batch: int, agents: int = 5, planes: int = 1, device="cuda", warp_flag=1,
                 inference="activated", training_pass_only=False, only_v2i=False, has_query=True, layer: int = 3):
        super().__init__(batch * agents, planes, device)
        prec, planes = self.prec, self.prec.planes
        if layer not in (2, 3, 4):
            raise ops.V2XError("When2com communicates at layer 2, 3 or 4 (When2com.py:167-190)")
        if layer != 3 and inference == "argmax_test" and not training_pass_only:
            raise ops.V2XError("argmax_test only exists for layer 3 in the reference (When2com.py:289-291)")
        ops.require_gpu()
        dev = self.device
        self.batch, self.agents = batch, agents
        f32 = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.enc_w = BackboneWeights(sd, "u_encoder.", prec, dev, encoder=True, decoder=False)
        self.dec_w = BackboneWeights(sd, "decoder.", prec, dev, encoder=False, decoder=True)
        # the policy branch feeds DISCRETE gates (p > 0.2 / argmax): it always runs at full split precision
        self.pol_w = BackboneWeights(sd, "query_key_net.lidar_encoder.", prec if planes == 1 else "fp16x3", dev,
                                     encoder=True, decoder=False)
        self.head_w = HeadWeights(sd, prec, dev)
        self.pol_convs = []
        for name, stride, cin in (("conv1", 1, 512), ("conv2", 1, 512), ("conv3", 2, 256), ("conv4", 1, 256),
                                  ("conv5", 2, 256)):
            pre = "query_key_net.%s.cbr_unit." % name
            self.pol_convs.append(ops.pack_conv(sd[pre + "0.weight"], sd[pre + "0.bias"], _bn(sd, pre + "1"), cins=[cin],
                                                stride=stride, planes=planes, device=dev))
        nets_used = ("key_net", "query_net") if has_query else ("key_net",)
        self.mlp = {net: [(f32("%s.fc.%d.weight" % (net, i)), f32("%s.fc.%d.bias" % (net, i))) for i in (0, 2, 4)]
                    for net in nets_used}
        self.att_w, self.att_b = f32("attention_net.linear.weight"), f32("attention_net.linear.bias")
        self.trans = torch.zeros((batch, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch, agents), agents, dtype=torch.int64, device=dev)
        trans, na, n = self.trans, self.num_agent, self.n

        x_in = self.build_input()
        # layer 4: the fuse reads the plain 16x16 x_4; the decoder gets each fused map upsampled (DetPlan.upsample2)
        x0, x1, x2, x3, x4u = self.build_encoder(self.enc_w, x_in, upsample_x4=(layer != 4))
        # ---- policy branch: second encoder's x_4 -> 5 convs -> [N,4,4,256] -> key / query MLPs ----
        t = self.build_encoder(self.pol_w, x_in, tag="pol_", upsample_x4=False)[4]
        for i, pc in enumerate(self.pol_convs):
            t = self.conv(pc, [t], "pol_c%d" % (i + 1))
        qk = t
        # has_query=False: query = ones(batch, 1, query_size) for every agent (When2com.py:241-245)
        feats = {"query_net": torch.ones((n, self.att_w.shape[1]), dtype=torch.float32, device=dev)}
        for net in nets_used:
            (w0, b0), (w1, b1), (w2, b2) = self.mlp[net]
            h0 = torch.empty((n, w0.shape[0]), dtype=torch.float32, device=dev)
            h1 = torch.empty((n, w1.shape[0]), dtype=torch.float32, device=dev)
            h2 = torch.empty((n, w2.shape[0]), dtype=torch.float32, device=dev)
            self.add(lambda w0=w0, b0=b0, h0=h0: ops.linear(qk, w0, b0, relu=True, out=h0, act_input=True))
            self.add(lambda w1=w1, b1=b1, h0=h0, h1=h1: ops.linear(h0, w1, b1, relu=True, out=h1))
            self.add(lambda w2=w2, b2=b2, h1=h1, h2=h2: ops.linear(h1, w2, b2, relu=False, out=h2))
            feats[net] = h2
        self.keys, self.querys = feats["key_net"], feats["query_net"]
        self.attn = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
        self.coef = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
        gate = ops.GATE_MODES[inference]
        keys, querys, attn, coef, aw, ab = self.keys, self.querys, self.attn, self.coef, self.att_w, self.att_b
        self.add(lambda: ops.attn_scores(keys, querys, aw, ab, batch, agents, gate, attn=attn, coef=coef))
        # ---- pass 1: softmax-weighted fuse of the communicated layer -> decoder ----
        xs = [x0, x1, x2, x3, x4u]
        xl = xs[layer]
        cl, hl, wl = xl.shape[-1], xl.shape[2], xl.shape[3]
        fuse1 = self.act("fuse1", hl, wl, cl)
        self.add(lambda: ops.warp_gated(xl, trans, na, attn, batch, agents, warp_flag=warp_flag, only_v2i=only_v2i,
                                        out=fuse1))
        xs[layer] = fuse1 if layer != 4 else self.upsample2(fuse1, "fuse1u")
        x8 = self.build_decoder(self.dec_w, xs[0], xs[1], xs[2], xs[3], xs[4])
        if not training_pass_only and inference != "softmax":
            fuse2 = self.act("fuse2", hl, wl, cl)
            self.add(lambda: ops.warp_gated(xl, trans, na, coef, batch, agents, warp_flag=warp_flag, only_v2i=only_v2i,
                                            out=fuse2))
            # the layer-0 skip of the second pass is the first pass's output (:266-270)
            xs[0], xs[layer] = x8, (fuse2 if layer != 4 else self.upsample2(fuse2, "fuse2u"))
            x8 = self.build_decoder(self.dec_w, xs[0], xs[1], xs[2], xs[3], xs[4], tag="p2_")
        self.build_heads(self.head_w, x8)

    def forward(self, bevs, trans_matrices, num_agent_tensor):
        self.bev_in.copy_(bevs.reshape(self.bev_in.shape), non_blocking=True)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)
        self.run()
        return self.result()


class PeerStep:
    """One kernel of the device-side exchange (sharding.PeerRegion): it synchronises with the other ranks, so tools that
    replay a subset of a plan's launches on one rank (bench.roofline_of, profilers) must skip it (``collective``)."""
    collective = True
    flops = 0.0

    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        self.fn()


class V2VNetDetShardedPlan(DetPlan):
    """det V2VNet forward, unit-sharded across ``world`` ranks (one process per GPU).

    This rank runs encoder / GRU / decoder / heads for its slice of the agent-major units and exchanges only the
    layer-3 maps: one NCCL all-gather of x_3, issued asynchronously right after conv3_2 so that conv4_1 / conv4_2
    overlap it.  The forward is three CUDA graphs (-> x_3 | x_4 branch | fuse + decoder + heads) around the collective."""

    def __init__(self, sd, batch_total: int, agents: int, rank: int, world: int, gnn_iter: int = 3, planes: int = 1,
                 device="cuda", group=None, only_v2i=False, exchange=None):
        from . import sharding
        # "allgather": one ncclAllGather of every unit's x_3; "neighbours": point-to-point exchange of just the maps of
        # the other agents of this rank's scenes (sharding.exchange_neighbour_units)
        self.exchange = exchange or os.environ.get("V2X_EXCHANGE", "allgather")
        assert self.exchange in ("allgather", "neighbours", "push")
        self.offset, n_loc = sharding.unit_range(batch_total * agents, rank, world)
        super().__init__(n_loc, planes, device)
        prec, planes = self.prec, self.prec.planes
        ops.require_gpu()
        self.sharding, self.group, self.world = sharding, group, world
        self.batch, self.agents, self.gnn_iter = batch_total, agents, gnn_iter
        dev = self.device
        self.enc_w = BackboneWeights(sd, "u_encoder.", prec, dev, encoder=True, decoder=False)
        self.dec_w = BackboneWeights(sd, "decoder.", prec, dev, encoder=False, decoder=True)
        self.head_w = HeadWeights(sd, prec, dev)
        self.gru_h, self.gru_m = ops.pack_gru_split(sd["convgru.weight_ih_l0"], sd["convgru.bias_ih_l0"],
                                                    sd["convgru.bias_hh_l0"], planes=planes, device=dev,
                                                    pre_act=GRU_PRE_ACT and planes == 1, mmas=prec.mmas("gru"))
        self.trans = torch.zeros((batch_total, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch_total, agents), agents, dtype=torch.int64, device=dev)
        trans, na, off = self.trans, self.num_agent, self.offset

        x_in = self.build_input()
        c = self.enc_w.c
        t = self.conv(c["conv_pre_1"], [x_in], "x0a")
        x0 = self.conv(c["conv_pre_2"], [t], "x0")
        t = self.conv(c["conv1_1"], [x0], "x1a")
        t = self.conv(c["conv1_2"], [t], "x1b")
        x1 = self.conv(c["conv3d_1"], [t], "x1")
        t = self.conv(c["conv2_1"], [x1], "x2a")
        t = self.conv(c["conv2_2"], [t], "x2b")
        x2 = self.conv(c["conv3d_2"], [t], "x2")
        t = self.conv(c["conv3_1"], [x2], "x3a")
        x3 = x3_raw = self.conv(c["conv3_2"], [t], "x3")
        if "com_compresser" in c:
            t = self.conv(c["com_compresser"], [x3_raw], "x3c")
            x3 = self.conv(c["com_decompresser"], [t], "x3d")
        self.stage_a = len(self.launches)            # ---- x_3 ready: the all-gather starts here
        t = self.conv(c["conv4_1"], [x3_raw], "x4a")
        x4u = self.conv(c["conv4_2"], [t], "x4u", upsample2x=True)
        self.stage_b = len(self.launches)            # ---- x_4 branch done: wait for the gather
        c3 = x3.shape[-1]
        self.x3_local = x3
        self.peer = None
        if self.exchange == "push":
            # device-side exchange over NVLink peer memory: x3_all lives in this rank's peer-visible region and every rank
            # stores its own maps straight into it (sharding.PeerRegion, csrc/peer_kernels.cu)
            x3_bytes = planes * n_loc * world * 32 * 32 * c3 * 2
            self.peer = sharding.PeerRegion(x3_bytes, rank, world, group=group, device=dev)
            self.x3_all = self.peer.payload((planes, n_loc * world, 32, 32, c3), ops.act_dtype(planes))
        else:
            self.x3_all = ops.empty_act(planes, n_loc * world, 32, 32, c3, dev)
            self.x3_all.zero_()
        # Planes that cross the wire.  The gathered maps only feed the neighbour mean, and in the "mixed" precision the
        # ConvGRU reads that mean through ONE tensor-core pass (its fp16 hi plane, 11 bits): sending the neighbours' lo
        # planes would move twice the bytes for bits the consumer rounds away.  So only the hi plane of REMOTE units is
        # exchanged (their lo plane stays zero); this rank's own units keep both planes.  Measured at 8 GPUs the all-gather
        # of both planes (294 MB received per rank) left 0.33 ms of a 4.8 ms step exposed.  V2X_EXCHANGE_PLANES overrides.
        self.exchange_planes = int(os.environ.get("V2X_EXCHANGE_PLANES", 0)) or (1 if (planes == 2 and prec.mmas("gru") == 1) else planes)
        x3_all = self.x3_all
        mean = self.act("mean", 32, 32, c3)
        if self.peer is not None:
            # begin (peers have consumed the previous step) -> push own maps into every region + publish -> wait for all
            # ranks' maps; the x_4 branch [stage_a, stage_b) runs beside all of it on the side stream of the ONE graph
            peer, ep = self.peer, self.exchange_planes
            dst_planes = [planes if r == rank else ep for r in range(world)]
            unit_elems = 32 * 32 * c3
            self.add(PeerStep(peer.begin))
            self.add(PeerStep(lambda: peer.push(x3, off * unit_elems, n_loc * world * unit_elems, dst_planes)))
            self.add(PeerStep(peer.wait))
        self.add(lambda: ops.warp_mean(x3_all, trans, na, batch_total, agents, include_self=False, only_v2i=only_v2i,
                                       out=mean, unit_offset=off, unit_count=n_loc))
        if self.peer is not None:
            self.add(PeerStep(self.peer.done))
        h = self.build_gru_rounds(x3, mean, gnn_iter, batch_total, agents, off)
        if self.peer is not None:
            self.side_lo, self.side_hi, self.side_join = self.stage_a, self.stage_b, len(self.launches)
        x8 = self.build_decoder(self.dec_w, x0, x1, x2, h, x4u)
        self.build_heads(self.head_w, x8)
        self.graphs = None

    def _segments(self):
        return (self.launches[:self.stage_a], self.launches[self.stage_a:self.stage_b], self.launches[self.stage_b:])

    def run(self):
        if self.peer is not None:     # one graph, no host-issued collective
            return DetPlan.run(self)
        seg = self._segments()
        for i in range(3):
            if i == 1:   # x_3 is final: exchange it while the x_4 branch runs
                ep = self.exchange_planes
                if self.exchange == "neighbours":
                    works = self.sharding.exchange_neighbour_units(self.x3_local[:ep], self.x3_all[:ep], self.batch,
                                                                   self.agents, group=self.group)
                else:
                    _, works = self.sharding.all_gather_units(self.x3_local[:ep], out=self.x3_all[:ep], group=self.group,
                                                              async_op=True)
                if ep < self.planes:   # own units keep their lo plane
                    self.x3_all[ep:, self.offset:self.offset + self.n].copy_(self.x3_local[ep:], non_blocking=True)
            if i == 2:
                for w in works:
                    w.wait()   # stream-level wait: the fuse kernels queue behind the collective
            if self.graphs is not None:
                self.graphs[i].replay()
            else:
                for l in seg[i]:
                    l()

    def capture(self):
        if self.peer is not None:
            if self.world > 1:
                import torch.distributed as dist
                dist.barrier(group=self.group)   # line the ranks up: the waits of the warm-up steps have a time limit
            DetPlan.capture(self)     # two eager warm-up steps on every rank (they exchange like real ones), then one graph
            self.peer.check()
            return
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                for l in self.launches:
                    l()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graphs = []
        for seg in self._segments():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for l in seg:
                    l()
            graphs.append(g)
        self.graphs = graphs

    def close(self):
        """Release the peer-memory region of the "push" exchange (collective across the ranks; no-op otherwise)."""
        if self.peer is not None:
            self.graph = None
            self.x3_all = None
            self.launches = []
            self.peer.close()
            self.peer = None

    def set_inputs(self, bevs_local, trans_matrices, num_agent_tensor):
        """bevs_local: this rank's unit slice [n_loc,1,256,256,13]; poses / agent counts of ALL scenes."""
        self.bev_in.copy_(bevs_local.reshape(self.bev_in.shape), non_blocking=True)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)

    def forward(self, bevs_local, trans_matrices, num_agent_tensor):
        self.set_inputs(bevs_local, trans_matrices, num_agent_tensor)
        self.run()
        return self.result()


class When2comDetShardedPlan(DetPlan):
    """det When2com / who2com forward (eval), unit-sharded across ``world`` ranks ("When2com additionally
    all-gathers keys [units,1024] and queries [units,32]").

    Every rank runs both encoders, the key / query MLPs, the two decoder passes and the heads for its contiguous slice of
    the agent-major units.  The exchange step: one NCCL all-gather of the keys and of the queries (rows = global units; a
    few hundred KB) -- the 5x5 attention of every scene needs the keys of all its agents -- after which each rank evaluates
    the (tiny) attention kernel for all scenes and fuses its own targets.  With warp_flag = 1 the fuse of target q reads
    only q's OWN map warped into the other agents' frames (the reference's val_mat[b,k,q] pairing), so no
    feature maps cross the wire at all; with warp_flag = 0 it reads every agent's map and x_3 is all-gathered too.
    The forward is two CUDA graphs around the collectives."""

    def __init__(self, sd, batch_total: int, agents: int, rank: int, world: int, planes=1, device="cuda", group=None,
                 warp_flag=1, inference="activated", only_v2i=False):
        from . import sharding
        self.offset, n_loc = sharding.unit_range(batch_total * agents, rank, world)
        super().__init__(n_loc, planes, device)
        prec, planes = self.prec, self.prec.planes
        ops.require_gpu()
        self.sharding, self.group, self.world = sharding, group, world
        self.batch, self.agents, self.warp_flag = batch_total, agents, warp_flag
        dev = self.device
        f32 = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.enc_w = BackboneWeights(sd, "u_encoder.", prec, dev, encoder=True, decoder=False)
        self.dec_w = BackboneWeights(sd, "decoder.", prec, dev, encoder=False, decoder=True)
        self.pol_w = BackboneWeights(sd, "query_key_net.lidar_encoder.", prec if planes == 1 else "fp16x3", dev,
                                     encoder=True, decoder=False)
        self.head_w = HeadWeights(sd, prec, dev)
        self.pol_convs = []
        for name, stride, cin in (("conv1", 1, 512), ("conv2", 1, 512), ("conv3", 2, 256), ("conv4", 1, 256),
                                  ("conv5", 2, 256)):
            pre = "query_key_net.%s.cbr_unit." % name
            self.pol_convs.append(ops.pack_conv(sd[pre + "0.weight"], sd[pre + "0.bias"], _bn(sd, pre + "1"), cins=[cin],
                                                stride=stride, planes=planes, device=dev))
        self.mlp = {net: [(f32("%s.fc.%d.weight" % (net, i)), f32("%s.fc.%d.bias" % (net, i))) for i in (0, 2, 4)]
                    for net in ("key_net", "query_net")}
        self.att_w, self.att_b = f32("attention_net.linear.weight"), f32("attention_net.linear.bias")
        self.trans = torch.zeros((batch_total, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch_total, agents), agents, dtype=torch.int64, device=dev)
        trans, na, off = self.trans, self.num_agent, self.offset
        units = batch_total * agents

        x_in = self.build_input()
        x0, x1, x2, x3, x4u = self.build_encoder(self.enc_w, x_in)
        t = self.build_encoder(self.pol_w, x_in, tag="pol_", upsample_x4=False)[4]
        for i, pc in enumerate(self.pol_convs):
            t = self.conv(pc, [t], "pol_c%d" % (i + 1))
        qk = t
        feats = {}
        for net in ("key_net", "query_net"):
            (w0, b0), (w1, b1), (w2, b2) = self.mlp[net]
            h0 = torch.empty((n_loc, w0.shape[0]), dtype=torch.float32, device=dev)
            h1 = torch.empty((n_loc, w1.shape[0]), dtype=torch.float32, device=dev)
            h2 = torch.empty((n_loc, w2.shape[0]), dtype=torch.float32, device=dev)
            self.add(lambda w0=w0, b0=b0, h0=h0: ops.linear(qk, w0, b0, relu=True, out=h0, act_input=True))
            self.add(lambda w1=w1, b1=b1, h0=h0, h1=h1: ops.linear(h0, w1, b1, relu=True, out=h1))
            self.add(lambda w2=w2, b2=b2, h1=h1, h2=h2: ops.linear(h1, w2, b2, relu=False, out=h2))
            feats[net] = h2
        self.keys_local, self.querys_local = feats["key_net"], feats["query_net"]
        self.stage_a = len(self.launches)            # ---- keys / queries (and x_3) ready: the exchange happens here
        self.keys = torch.empty((units, self.keys_local.shape[1]), dtype=torch.float32, device=dev)
        self.querys = torch.empty((units, self.querys_local.shape[1]), dtype=torch.float32, device=dev)
        self.x3_local = x3
        c3 = x3.shape[-1]
        x_src, x_off = x3, off
        if not warp_flag:
            self.x3_all = ops.empty_act(planes, units, 32, 32, c3, dev)
            x_src, x_off = self.x3_all, 0
        self.attn = torch.empty((batch_total, agents, agents), dtype=torch.float32, device=dev)
        self.coef = torch.empty((batch_total, agents, agents), dtype=torch.float32, device=dev)
        gate = ops.GATE_MODES[inference]
        keys, querys, attn, coef, aw, ab = self.keys, self.querys, self.attn, self.coef, self.att_w, self.att_b
        self.add(lambda: ops.attn_scores(keys, querys, aw, ab, batch_total, agents, gate, attn=attn, coef=coef))
        fuse1 = self.act("fuse1", 32, 32, c3)
        self.add(lambda: ops.warp_gated(x_src, trans, na, attn, batch_total, agents, warp_flag=warp_flag, only_v2i=only_v2i,
                                        out=fuse1, unit_offset=off, unit_count=n_loc, x_unit_offset=x_off))
        x8 = self.build_decoder(self.dec_w, x0, x1, x2, fuse1, x4u)
        if inference != "softmax":
            fuse2 = self.act("fuse2", 32, 32, c3)
            self.add(lambda: ops.warp_gated(x_src, trans, na, coef, batch_total, agents, warp_flag=warp_flag,
                                            only_v2i=only_v2i, out=fuse2, unit_offset=off, unit_count=n_loc,
                                            x_unit_offset=x_off))
            x8 = self.build_decoder(self.dec_w, x8, x1, x2, fuse2, x4u, tag="p2_")
        self.build_heads(self.head_w, x8)
        self.graphs = None

    def _segments(self):
        return (self.launches[:self.stage_a], self.launches[self.stage_a:])

    def _exchange(self):
        self.sharding.all_gather_units(self.keys_local, out=self.keys, group=self.group)
        self.sharding.all_gather_units(self.querys_local, out=self.querys, group=self.group)
        if not self.warp_flag:
            self.sharding.all_gather_units(self.x3_local, out=self.x3_all, group=self.group)

    def run(self):
        seg = self._segments()
        for i in range(2):
            if i == 1:
                self._exchange()      # stream-ordered NCCL collectives between the two graphs
            if self.graphs is not None:
                self.graphs[i].replay()
            else:
                for l in seg[i]:
                    l()

    def capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                for l in self.launches:
                    l()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graphs = []
        for seg in self._segments():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for l in seg:
                    l()
            graphs.append(g)
        self.graphs = graphs

    def forward(self, bevs_local, trans_matrices, num_agent_tensor):
        """bevs_local: this rank's unit slice [n_loc,1,256,256,13]; poses / agent counts of ALL scenes."""
        self.bev_in.copy_(bevs_local.reshape(self.bev_in.shape), non_blocking=True)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)
        self.run()
        return self.result()


def _pair_mlp(sd, p, dev):
    """Layers 2..4 of PixelWeightedFusionSoftmax / AgentWeightedFusion (DiscoNet.py:136-155) with BN(eval) folded."""
    w2, b2 = ops.fold_bn_1x1(sd[p + "conv1_2.weight"], sd[p + "conv1_2.bias"], _bn(sd, p + "bn1_2"), dev)
    w3, b3 = ops.fold_bn_1x1(sd[p + "conv1_3.weight"], sd[p + "conv1_3.bias"], _bn(sd, p + "bn1_3"), dev)
    w4, b4 = ops.fold_bn_1x1(sd[p + "conv1_4.weight"], sd[p + "conv1_4.bias"], None, dev)
    return (w2, b2, w3, b3, w4.reshape(-1).contiguous(), b4)


class FuseStage:
    """The cross-agent fuse step of the FusionBase family on an agent-major layer map ``x`` (act [P, A*B, h, w, C]),
    shared by the det (FusionBase.py:23-75, DiscoNet.py:36-129) and seg (seg/FusionBase.py:25-84) plans.

    kind: "mean" | "sum" | "max" | "cat" | "agent" | "disco".  Adds its launches to ``plan`` and returns the fused act."""

    KINDS = ("mean", "sum", "max", "cat", "agent", "disco")
    PREFIX = {"cat_det": "_modulation_layer_3._", "cat_seg": "modulation_layer_3.", "agent": "agent_weighted_fusion.",
              "disco": "pixel_weighted_fusion."}

    def __init__(self, plan, kind, sd, x, trans, num_agent, batch, agents, *, only_v2i=False, seg=False):
        assert kind in self.KINDS
        dev, planes = plan.device, plan.planes
        _, n, h, w, c = x.shape
        na = num_agent
        if kind in ("mean", "sum", "max"):
            out = plan.act("fused", h, w, c)
            plan.add(lambda: ops.warp_reduce(x, trans, na, batch, agents, kind, only_v2i=only_v2i, out=out))
        elif kind == "cat":
            # mean over the list, cat([tg, mean]) -> 1x1 conv + BN + ReLU (CatFusion.py:22-26,37-41) == a two-source conv
            p = self.PREFIX["cat_seg" if seg else "cat_det"]
            self.pc = ops.pack_conv(sd[p + "conv1_1.weight"], sd[p + "conv1_1.bias"], _bn(sd, p + "bn1_1"), cins=[c, c],
                                    planes=planes, device=dev)
            mean = plan.act("fuse_mean", h, w, c)
            plan.add(lambda: ops.warp_reduce(x, trans, na, batch, agents, "mean", only_v2i=only_v2i, out=mean))
            out = plan.conv(self.pc, [x, mean], "fused")
            plan.add(lambda: ops.restore_absent(x, out, na, batch, agents))
        else:
            p = self.PREFIX[kind]
            self.pc = ops.pack_pair_conv1(sd[p + "conv1_1.weight"], sd[p + "conv1_1.bias"], _bn(sd, p + "bn1_1"),
                                          planes=planes, device=dev)
            self.mlp = mlp = _pair_mlp(sd, p, dev)
            q = plan.act("fuse_q", h, w, 256)
            plan.add(ConvLaunch(self.pc, [x], epilogue=EPI_ACT, relu=False, out0=q))
            self.scores = scores = torch.zeros((batch, agents, agents, h * w), dtype=torch.float32, device=dev)
            plan.add(lambda: ops.pair_score(q, trans, na, batch, agents, mlp, only_v2i=only_v2i, out=scores))
            out = plan.act("fused", h, w, c)
            if kind == "disco":
                plan.add(lambda: ops.warp_weighted(x, trans, na, scores, batch, agents, per_pixel=True,
                                                   only_v2i=only_v2i, out=out))
            else:
                # conv1_5: 32x32 "valid" conv over the H-flipped score map -> one scalar per pair
                # (AgentWiseWeightedFusion.py:65,74); scores are un-flipped here, so mirror the filter rows instead
                if (h, w) != (32, 32):
                    raise ops.V2XError("AgentWiseWeightedFusion's 32x32 conv1_5 only fits 32x32 maps (layer 3), as in "
                                       "the reference (AgentWiseWeightedFusion.py:65)")
                w5 = sd[p + "conv1_5.weight"].detach().to(device=dev, dtype=torch.float32).reshape(h, w)
                self.w5f = w5f = torch.flip(w5, (0,)).contiguous()
                self.b5 = b5 = sd[p + "conv1_5.bias"].detach().to(device=dev, dtype=torch.float32).contiguous()
                self.coef = coef = torch.zeros((batch, agents, agents), dtype=torch.float32, device=dev)
                plan.add(lambda: ops.agent_softmax(scores, w5f, b5, na, out=coef))
                plan.add(lambda: ops.warp_weighted(x, trans, na, coef, batch, agents, per_pixel=False,
                                                   only_v2i=only_v2i, out=out))
        self.out = out


class FusionDetPlan(DetPlan):
    """det intermediate-fusion baselines: encoder -> FuseStage(kind) at layer 3 -> decoder -> heads
    (FusionBase.py:23-75; DiscoNet.py:36-129)."""

    def __init__(self, sd, kind: str, batch: int, agents: int = 5, planes: int = 1, device="cuda", only_v2i=False,
                 layer: int = 3):
        super().__init__(batch * agents, planes, device)
        prec, planes = self.prec, self.prec.planes
        if layer not in (0, 1, 2, 3, 4):
            raise ops.V2XError("fusion models fuse at layer 0..4 (DetModelBase.py:71-92)")
        ops.require_gpu()
        self.batch, self.agents, self.kind = batch, agents, kind
        dev = self.device
        self.enc_w = BackboneWeights(sd, "u_encoder.", prec, dev, encoder=True, decoder=False)
        self.dec_w = BackboneWeights(sd, "decoder.", prec, dev, encoder=False, decoder=True)
        self.head_w = HeadWeights(sd, prec, dev)
        self.trans = torch.zeros((batch, agents, agents, 4, 4), dtype=torch.float64, device=dev)
        self.num_agent = torch.full((batch, agents), agents, dtype=torch.int64, device=dev)
        x_in = self.build_input()
        # layer 4: the fuse reads the plain 16x16 x_4; the decoder gets the fused map upsampled (DetPlan.upsample2)
        x0, x1, x2, x3, x4u = self.build_encoder(self.enc_w, x_in, upsample_x4=(layer != 4))
        xs = [x0, x1, x2, x3, x4u]
        self.fuse = FuseStage(self, kind, sd, xs[layer], self.trans, self.num_agent, batch, agents, only_v2i=only_v2i)
        self.fused = xs[layer] = self.fuse.out
        if layer == 4:
            xs[4] = self.upsample2(self.fused, "x4u")
        x8 = self.build_decoder(self.dec_w, xs[0], xs[1], xs[2], xs[3], xs[4])
        self.build_heads(self.head_w, x8)

    def forward(self, bevs, trans_matrices, num_agent_tensor):
        self.bev_in.copy_(bevs.reshape(self.bev_in.shape), non_blocking=True)
        self.trans.copy_(trans_matrices.reshape(self.trans.shape), non_blocking=True)
        self.num_agent.copy_(num_agent_tensor.reshape(self.num_agent.shape), non_blocking=True)
        self.run()
        return self.result()

    def kd_layers(self):
        """(x_8, x_7, x_6, x_5, fused) as fp32 NCHW tensors -- what kd_flag == 1 forwards return (FusionBase.py:72-73).
        x_5..x_7 only exist nearest-upsampled in the workspace; every 2x2 block holds one value."""
        f = ops.act_to_float
        return (f(self.ws["x8"]), f(self.ws["x7u"])[:, :, ::2, ::2].contiguous(), f(self.ws["x6u"])[:, :, ::2, ::2].contiguous(),
                f(self.ws["x5u"])[:, :, ::2, ::2].contiguous(), f(self.fused))
