"""TEST INFRASTRUCTURE ONLY -- fp32 torch-CPU restatement of the reference hot path.

A literal, functional (no nn.Module) restatement of coperception's detection forward,
driven by a reference-format ``state_dict``.  It deliberately keeps the reference's
structure -- the H flip around the fusion stage, the per-agent / per-round python loops,
the ConvGRU conv over the all-zero hidden state, warps recomputed in every GNN round --
because it doubles as the "reference CPU implementation" timed by ``bench.py``'s
``cpu_baseline`` leg (kind = "port").  Pinned against the live reference modules through
``tests/golden/*.npz`` (see ``oracle/gen_golden.py``); never imported by the product.

Shorthand: CP/ = /root/reference/coperception/coperception/
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default


_TRAIN = False      # set by ``training_mode()``: BatchNorm uses batch statistics and updates its running buffers
BN_MOMENTUM = 0.1   # nn.BatchNorm2d / BatchNorm3d default (the reference never overrides it)


class training_mode:
    """``with restate.training_mode(): ...`` -- the restatement behaves like the reference module after ``.train()``:
    every BatchNorm normalises with the statistics of the current batch (all A*B maps, padded agents included), updates
    ``running_mean`` / ``running_var`` in place with momentum 0.1 (unbiased variance) and bumps ``num_batches_tracked``
    (torch.nn.modules.batchnorm._BatchNorm.forward).  Nothing else on the path differs between train and eval
    (no dropout; ``p_com_outage`` defaults to 0).  Used as the oracle of the training / backward row, SURVEY 8(f1):
    gradients come from torch.autograd over these same functions."""

    def __enter__(self):
        global _TRAIN
        self._prev, _TRAIN = _TRAIN, True
        return self

    def __exit__(self, *exc):
        global _TRAIN
        _TRAIN = self._prev
        return False


def _bn(x, sd, name):
    if _TRAIN:
        if name + ".num_batches_tracked" in sd:
            sd[name + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                            sd[name + ".weight"], sd[name + ".bias"], True, BN_MOMENTUM, EPS)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, EPS)


def cbr(x, sd, p, conv, bn, stride=1):
    """conv3x3(pad 1) + BN(eval) + ReLU  (CP/models/det/backbone/Backbone.py:102-136)."""
    y = F.conv2d(x, sd[p + conv + ".weight"], sd[p + conv + ".bias"], stride=stride, padding=1)
    return F.relu(_bn(y, sd, p + bn))


def conv3d_111(x, sd, p, name):
    """Conv3D 1x1x1 + BN3d + ReLU over a length-1 sequence (Backbone.py:280-300) == 1x1 conv2d."""
    w = sd[p + name + ".conv3d.weight"]
    y = F.conv2d(x, w.reshape(w.shape[0], w.shape[1], 1, 1), sd[p + name + ".conv3d.bias"])
    return F.relu(_bn(y, sd, p + name + ".bn3d"))


def encode(bevs, sd, p="u_encoder.", compress_level=0):
    """Backbone.encode (Backbone.py:89-143).  bevs: [N,1,256,256,13] (the model permutes to
    [N,1,13,256,256] first, V2VNet.py:51)."""
    x = bevs.permute(0, 1, 4, 2, 3)
    # Backbone.py:101 casts to torch.float; following the weights' dtype instead is identical in fp32 and lets the
    # training oracle run in float64 (its fixtures are made from the live module under a float->double shim)
    x = x.reshape(-1, x.size(-3), x.size(-2), x.size(-1)).to(sd[p + "conv_pre_1.weight"].dtype)
    x = cbr(x, sd, p, "conv_pre_1", "bn_pre_1")
    x = cbr(x, sd, p, "conv_pre_2", "bn_pre_2")
    x_1 = cbr(x, sd, p, "conv1_1", "bn1_1", stride=2)
    x_1 = cbr(x_1, sd, p, "conv1_2", "bn1_2")
    x_1 = conv3d_111(x_1, sd, p, "conv3d_1")
    x_2 = cbr(x_1, sd, p, "conv2_1", "bn2_1", stride=2)
    x_2 = cbr(x_2, sd, p, "conv2_2", "bn2_2")
    x_2 = conv3d_111(x_2, sd, p, "conv3d_2")
    x_3 = cbr(x_2, sd, p, "conv3_1", "bn3_1", stride=2)
    x_3 = cbr(x_3, sd, p, "conv3_2", "bn3_2")
    x_4 = cbr(x_3, sd, p, "conv4_1", "bn4_1", stride=2)
    x_4 = cbr(x_4, sd, p, "conv4_2", "bn4_2")
    if compress_level > 0:  # Backbone.py:139-141
        y = F.conv2d(x_3, sd[p + "com_compresser.weight"], sd[p + "com_compresser.bias"])
        x_3 = F.relu(_bn(y, sd, p + "bn_compress"))
        y = F.conv2d(x_3, sd[p + "com_decompresser.weight"], sd[p + "com_decompresser.bias"])
        x_3 = F.relu(_bn(y, sd, p + "bn_decompress"))
    return [x, x_1, x_2, x_3, x_4]


def decode(x, x_1, x_2, x_3, x_4, sd, p="decoder.", kd_flag=False):
    """Backbone.decode (Backbone.py:145-242); nearest x2 upsample, cat((up, skip)), two CBRs.
    The permute / adaptive_max_pool3d blocks are identities for seq=1."""
    x_5 = cbr(torch.cat((F.interpolate(x_4, scale_factor=(2, 2)), x_3), 1), sd, p, "conv5_1", "bn5_1")
    x_5 = cbr(x_5, sd, p, "conv5_2", "bn5_2")
    x_6 = cbr(torch.cat((F.interpolate(x_5, scale_factor=(2, 2)), x_2), 1), sd, p, "conv6_1", "bn6_1")
    x_6 = cbr(x_6, sd, p, "conv6_2", "bn6_2")
    x_7 = cbr(torch.cat((F.interpolate(x_6, scale_factor=(2, 2)), x_1), 1), sd, p, "conv7_1", "bn7_1")
    x_7 = cbr(x_7, sd, p, "conv7_2", "bn7_2")
    x_8 = cbr(torch.cat((F.interpolate(x_7, scale_factor=(2, 2)), x), 1), sd, p, "conv8_1", "bn8_1")
    res = cbr(x_8, sd, p, "conv8_2", "bn8_2")
    return [res, x_7, x_6, x_5] if kd_flag else [res]


def heads(x, sd):
    """ClassificationHead / SingleRegressionHead + get_cls_loc_result (DetModelBase.py:226-351)."""
    n = x.shape[0]
    c = F.relu(_bn(F.conv2d(x, sd["classification.conv1.weight"], sd["classification.conv1.bias"], padding=1),
                   sd, "classification.bn1"))
    c = F.conv2d(c, sd["classification.conv2.weight"], sd["classification.conv2.bias"])
    cls = c.permute(0, 2, 3, 1).contiguous().view(n, -1, 2)
    r = F.relu(_bn(F.conv2d(x, sd["regression.box_prediction.0.weight"], sd["regression.box_prediction.0.bias"],
                            padding=1), sd, "regression.box_prediction.1"))
    r = F.conv2d(r, sd["regression.box_prediction.3.weight"], sd["regression.box_prediction.3.bias"])
    r = r.permute(0, 2, 3, 1).contiguous()
    loc = r.view(-1, r.size(1), r.size(2), 6, 1, 6)
    return {"loc": loc, "cls": cls}


def feature_transformation(local_com_mat, b, j, agent_idx, trans_matrices, size):
    """DetModelBase.feature_transformation (DetModelBase.py:139-169): warp agent j's (H-flipped)
    map into agent_idx's frame; theta = [R | -t * 4/128], affine_grid + grid_sample defaults
    (bilinear, zeros padding, align_corners=False)."""
    nb_agent = local_com_mat[b, j].unsqueeze(0)
    tfm = trans_matrices[b, j, agent_idx]
    M = torch.hstack((tfm[:2, :2], -tfm[:2, 3:4])).to(local_com_mat.dtype).unsqueeze(0)   # `.float()` in the reference
    mask = torch.tensor([[[1, 1, 4 / 128], [1, 1, 4 / 128]]], dtype=M.dtype)
    M = M * mask
    grid = F.affine_grid(M, size=torch.Size(size), align_corners=False)
    return F.grid_sample(nb_agent, grid, mode="bilinear", padding_mode="zeros", align_corners=False).squeeze(0)


def convgru_zero_hidden(x, sd, p="convgru."):
    """Conv2dGRU, one layer, one time step, hx=None -> zeros (module.py:171-211; GRUCell
    functional.py:84-105; same-padding conv functional.py:261-301).  x: [1,Cin,H,W]."""
    w_ih, w_hh = sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"]
    b_ih, b_hh = sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"]
    hidden = x.new_zeros(x.shape[0], w_hh.shape[1], x.shape[2], x.shape[3])
    gi = F.conv2d(F.pad(x, (1, 1, 1, 1)), w_ih, b_ih)
    gh = F.conv2d(F.pad(hidden, (1, 1, 1, 1)), w_hh, b_hh)  # executed by the reference although hidden == 0
    i_r, i_i, i_n = gi.chunk(3, 1)
    h_r, h_i, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_i + h_i)
    n = torch.tanh(i_n + r * h_n)
    return n + z * (hidden - n)


def v2vnet_fuse(x_3, trans_matrices, num_agent_tensor, sd, batch_size, agent_num=5, gnn_iter=3,
                return_mean=False):
    """The GNN block of det V2VNet.forward (V2VNet.py:55-112) incl. the flips
    (DetModelBase.py:71-92, 53-69).  x_3: [A*B,C,32,32] agent-major."""
    c, h, w = x_3.shape[1:]
    size = (1, c, h, w)
    feat = torch.flip(x_3, (2,))
    local = torch.stack([feat[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)  # [B,A,C,H,W]
    upd = local.clone()
    means = torch.zeros_like(local)
    for b in range(batch_size):
        na = int(num_agent_tensor[b, 0])
        agent_feat = [local[b, i] for i in range(agent_num)]
        for _ in range(gnn_iter):
            new = []
            for i in range(na):
                nb = [feature_transformation(local, b, j, i, trans_matrices, size) for j in range(na) if j != i]
                mean = torch.mean(torch.stack(nb), dim=0)  # self excluded (V2VNet.py:96-98)
                means[b, i] = mean
                cat = torch.cat([agent_feat[i], mean], 0).unsqueeze(0)
                new.append(convgru_zero_hidden(cat, sd).squeeze(0))
            agent_feat = new
        for k in range(na):
            upd[b, k] = agent_feat[k]
    out = torch.flip(torch.cat([upd[:, i] for i in range(agent_num)], 0), (2,))
    if return_mean:
        m = torch.flip(torch.cat([means[:, i] for i in range(agent_num)], 0), (2,))
        return out, m
    return out


def v2vnet_det_forward(bevs, trans_matrices, num_agent_tensor, sd, batch_size=1, agent_num=5, gnn_iter=3,
                       compress_level=0, stages=False, layer=3):
    """det V2VNet.forward (CP/models/det/V2VNet.py:47-120); ``layer`` = the communicated encoder layer
    (DetModelBase.get_feature_maps_and_size / get_decoded_layers, :71-92, :211-224)."""
    enc = encode(bevs, sd, "u_encoder.", compress_level)
    fused = v2vnet_fuse(enc[layer], trans_matrices, num_agent_tensor, sd, batch_size, agent_num, gnn_iter)
    dec_in = list(enc)
    dec_in[layer] = fused
    x8 = decode(*dec_in, sd, "decoder.")[0]
    res = heads(x8, sd)
    if stages:
        res = dict(res, enc=enc, fused=fused, x8=x8)
    return res


def fafnet_forward(bevs, sd, compress_level=0, stages=False):
    """FaFNet.forward with kd_flag=0 semantics (FaFNet.py:28-39, Backbone.py:245-257)."""
    enc = encode(bevs, sd, "stpn.", compress_level)
    dec = decode(*enc, sd, "stpn.", kd_flag=True)
    res = heads(dec[0], sd)
    if stages:
        res = dict(res, enc=enc, dec=dec)
    return res


# ---------------------------------------------------------------------------------------------
# When2com / who2com detection (CP/models/det/When2com.py)
# ---------------------------------------------------------------------------------------------
def policy_net4(bevs, sd, p="query_key_net."):
    """PolicyNet4.forward (When2com.py:353-359): a second LidarEncoder (its x_4) + five conv+BN+ReLU."""
    x = encode(bevs, sd, p + "lidar_encoder.")[4]
    for name, stride in (("conv1", 1), ("conv2", 1), ("conv3", 2), ("conv4", 1), ("conv5", 2)):
        x = cbr(x, sd, p + name + ".", "cbr_unit.0", "cbr_unit.1", stride=stride)
    return x


def km_generator(maps, sd, p):
    """KmGenerator.forward (When2com.py:428-430): view(-1, 256*4*4) -> Linear/ReLU/Linear/ReLU/Linear."""
    x = maps.reshape(-1, 256 * 4 * 4)
    x = F.relu(F.linear(x, sd[p + "fc.0.weight"], sd[p + "fc.0.bias"]))
    x = F.relu(F.linear(x, sd[p + "fc.2.weight"], sd[p + "fc.2.bias"]))
    return F.linear(x, sd[p + "fc.4.weight"], sd[p + "fc.4.bias"])


def when2com_det_forward(bevs, trans_matrices, num_agent_tensor, sd, batch_size=1, agent_num=5, warp_flag=1,
                         inference="activated", training=False, only_v2i=False, stages=False, has_query=True, layer=3):
    """det When2com.forward (When2com.py:150-332), MO_flag=True.  Eval-mode semantics: ``inference`` in {"softmax",
    "activated", "argmax_test"}; training=True stops after the first decoder pass.  ``has_query=False``: every agent's
    query is a vector of ones (When2com.py:241-245).  ``layer`` in {2, 3, 4}: the communicated encoder layer (:167-190; the
    reference's argmax_test branch only exists for layer 3, :289-291).  ``sparse`` is not a parameter: the reference hands
    it to the attention module, which never reads it (:374-412)."""
    enc = encode(bevs, sd, "u_encoder.")
    x, x_1, x_2, x_3, x_4 = enc
    assert layer in (2, 3, 4) and not (layer != 3 and inference == "argmax_test" and not training)
    c, h, w = enc[layer].shape[1:]
    size = (1, c, h, w)
    feat = torch.flip(enc[layer], (2,))
    local = torch.stack([feat[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)  # [B,A,C,H,W]
    if warp_flag == 1:
        val_mat = torch.zeros(batch_size, agent_num, agent_num, c, h, w, dtype=local.dtype)
        for b in range(batch_size):
            na = int(num_agent_tensor[b, 0])
            for i in range(na):
                for j in range(na):
                    if j == i:
                        val_mat[b, i, j] = local[b, i]
                    else:
                        if only_v2i and i != 0 and j != 0:
                            continue
                        val_mat[b, i, j] = feature_transformation(local, b, j, i, trans_matrices, size)
    else:
        val_mat = local
    qk = policy_net4(bevs, sd)
    keys = km_generator(qk, sd, "key_net.")
    querys = km_generator(qk, sd, "query_net.") if has_query else keys.new_ones(
        (keys.shape[0], sd["attention_net.linear.weight"].shape[1]))
    key_mat = torch.stack([keys[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)      # [B,A,1024]
    query_mat = torch.stack([querys[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)  # [B,A,32]
    # MIMOGeneralDotProductAttention.forward (When2com.py:374-412)
    query = F.linear(query_mat, sd["attention_net.linear.weight"], sd["attention_net.linear.bias"])
    attn = torch.softmax(torch.bmm(key_mat, query.transpose(2, 1)), dim=1)  # [B, key, query]

    def weighted(coef):
        ce = coef.view(batch_size, agent_num, agent_num, 1, 1, 1)
        v = val_mat if warp_flag == 1 else val_mat.unsqueeze(2).expand(-1, -1, agent_num, -1, -1, -1)
        return (ce * v).sum(1)  # [B, query, C, H, W]

    def to_batch(f):  # agents_to_batch (DetModelBase.py:53-69): agent-major + flip back
        return torch.flip(torch.cat([f[:, i] for i in range(agent_num)], 0), (2,))

    def dec(x0, fused):   # the fused map replaces the communicated layer (:262-270)
        skips = [x0, x_1, x_2, x_3, x_4]
        skips[layer] = fused
        return decode(*skips, sd, "decoder.")[0]

    fuse1 = to_batch(weighted(attn))
    x_dec = dec(x, fuse1)
    prob = attn + torch.eye(agent_num).view(1, agent_num, agent_num) * 0.001
    fuse2 = None
    if not training:
        if inference == "softmax":
            pass
        elif inference in ("activated", "argmax_test"):
            if inference == "activated":
                coef = prob * (prob > 0.2).float()                       # activated_select (:125-148)
            else:
                coef = F.one_hot(prob.max(dim=1)[1], num_classes=agent_num).float().transpose(1, 2)  # argmax_select (:94-123)
            fuse2 = to_batch(weighted(coef))
            # NOTE: the layer-0 skip input of the second pass is the OUTPUT of the first pass (`x` was overwritten, :266-270)
            x_dec = dec(x_dec, fuse2)
        else:
            raise ValueError("Incorrect inference mode")
    res = heads(x_dec, sd)
    if stages:
        res = dict(res, enc=enc, qk=qk, keys=keys, querys=querys, attn=attn, fuse1=fuse1, fuse2=fuse2, x8=x_dec)
    return res


# ---------------------------------------------------------------------------------------------
# Segmentation models (CP/models/seg/)
# ---------------------------------------------------------------------------------------------
def double_conv(x, sd, p):
    """DoubleConv (SegModelBase.py:91-106): (conv3x3 + BN + ReLU) x 2; keys p+{0,1,3,4}."""
    x = F.relu(_bn(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1), sd, p + "1"))
    return F.relu(_bn(F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], padding=1), sd, p + "4"))


def seg_down(x, sd, p):
    """Down (SegModelBase.py:109-118): MaxPool2d(2) + DoubleConv."""
    return double_conv(F.max_pool2d(x, 2), sd, p + "maxpool_conv.1.double_conv.")


def seg_up(x1, x2, sd, p):
    """Up (SegModelBase.py:121-142): bilinear x2 (align_corners=True), cat([skip, up]), DoubleConv."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    return double_conv(torch.cat([x2, x1], dim=1), sd, p + "conv.double_conv.")


def seg_encode(x, sd, p=""):
    x = x.to(sd[p + "inc.double_conv.0.weight"].dtype)   # fp32 in the reference; float64 for the training-oracle pin
    x1 = double_conv(x, sd, p + "inc.double_conv.")
    x2 = seg_down(x1, sd, p + "down1.")
    x3 = seg_down(x2, sd, p + "down2.")
    x4 = seg_down(x3, sd, p + "down3.")
    if p + "com_compresser.weight" in sd:   # compress_level > 0 (UNet.py:30-32, seg/FusionBase.py:31-33, ...)
        x4 = F.relu(_bn(F.conv2d(x4, sd[p + "com_compresser.weight"], sd[p + "com_compresser.bias"]), sd, p + "bn_compress"))
        x4 = F.relu(_bn(F.conv2d(x4, sd[p + "com_decompresser.weight"], sd[p + "com_decompresser.bias"]), sd,
                        p + "bn_decompress"))
    return x1, x2, x3, x4


def seg_decode(feat, x1, x2, x3, sd):
    """down4 / up1..4 / outc on the (fused) layer-4 map (UNet.py:34-40, seg/V2VNet.py:85-91)."""
    x5 = seg_down(feat, sd, "down4.")
    x = seg_up(x5, feat, sd, "up1.")
    x = seg_up(x, x3, sd, "up2.")
    x = seg_up(x, x2, sd, "up3.")
    x = seg_up(x, x1, sd, "up4.")
    return F.conv2d(x, sd["outc.conv.weight"], sd["outc.conv.bias"])


def seg_unet_forward(x, sd):
    """seg UNet.forward (CP/models/seg/UNet.py:24-44), kd_flag=False."""
    x1, x2, x3, x4 = seg_encode(x, sd)
    return seg_decode(x4, x1, x2, x3, sd)


def seg_v2vnet_forward(x, trans_matrices, num_agent_tensor, sd, agent_num=5, only_v2i=False, stages=False):
    """seg V2VNet.forward (CP/models/seg/V2VNet.py:25-92): one GNN round, neighbour mean INCLUDES self (:55-56,74)."""
    x1, x2, x3, x4 = seg_encode(x, sd)
    c, h, w = x4.shape[1:]
    size = (1, c, h, w)
    batch_size = x.size(0) // agent_num
    feat = torch.flip(x4, (2,))
    local = torch.stack([feat[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)
    upd = local.clone()
    for b in range(batch_size):
        na = int(num_agent_tensor[b, 0])
        new = []
        for i in range(na):
            nb = [local[b, i]]
            for j in range(na):
                if j != i:
                    if only_v2i and i != 0 and j != 0:
                        continue
                    nb.append(feature_transformation(local, b, j, i, trans_matrices, size))
            mean = torch.mean(torch.stack(nb), dim=0)
            cat = torch.cat([local[b, i], mean], 0).unsqueeze(0)
            new.append(convgru_zero_hidden(cat, sd).squeeze(0))
        for k in range(na):
            upd[b, k] = new[k]
    fused = torch.flip(torch.cat([upd[:, i] for i in range(agent_num)], 0), (2,))
    logits = seg_decode(fused, x1, x2, x3, sd)
    return dict(logits=logits, x4=x4, fused=fused) if stages else logits


def seg_policy_net4(x, sd, p="query_key_net."):
    """seg PolicyNet4.forward (When2Com_UNet.py:331-339): own inc/down1..3, then five conv+BN+ReLU -> [N,256,8,8]."""
    t = seg_encode(x, sd, p)[3]
    for name, stride in (("conv1", 1), ("conv2", 1), ("conv3", 2), ("conv4", 1), ("conv5", 2)):
        t = cbr(t, sd, p + name + ".", "cbr_unit.0", "cbr_unit.1", stride=stride)
    return t


def seg_when2com_forward(x, trans_matrices, num_agent_tensor, sd, agent_num=5, warp_flag=1, inference="activated",
                         training=False, only_v2i=False, stages=False, has_query=True):
    """seg When2Com_UNet.forward (When2Com_UNet.py:144-307).  Reproduces the key/query row quirk: PolicyNet4 emits
    256*8*8 = 16384 features per map but KmGenerator views them as rows of 4096 (:381-392), and rows 0..A*B-1 of the
    resulting [4*A*B, .] matrices are taken as the agents' keys / queries (:207-226)."""
    x1, x2, x3, x4 = seg_encode(x, sd)
    c, h, w = x4.shape[1:]
    size = (1, c, h, w)
    batch_size = x.size(0) // agent_num
    feat = torch.flip(x4, (2,))
    local = torch.stack([feat[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)
    if warp_flag == 1:
        val_mat = torch.zeros(batch_size, agent_num, agent_num, c, h, w, dtype=local.dtype)
        for b in range(batch_size):
            na = int(num_agent_tensor[b, 0])
            for i in range(na):
                for j in range(na):
                    if j == i:
                        val_mat[b, i, j] = local[b, i]
                    else:
                        if only_v2i and i != 0 and j != 0:
                            continue
                        val_mat[b, i, j] = feature_transformation(local, b, j, i, trans_matrices, size)
    else:
        val_mat = local
    qk = seg_policy_net4(x, sd)
    keys = km_generator(qk, sd, "key_net.")      # [4*A*B, 1024]
    # has_query=False: every agent's query is a vector of ones (When2Com_UNet.py:219-225)
    querys = km_generator(qk, sd, "query_net.") if has_query else keys.new_ones(
        (keys.shape[0], sd["attention_net.linear.weight"].shape[1]))  # [4*A*B, 32]
    key_mat = torch.stack([keys[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)
    query_mat = torch.stack([querys[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)
    query = F.linear(query_mat, sd["attention_net.linear.weight"], sd["attention_net.linear.bias"])
    attn = torch.softmax(torch.bmm(key_mat, query.transpose(2, 1)), dim=1)
    prob = attn + torch.eye(agent_num).view(1, agent_num, agent_num) * 0.001
    if training or inference == "softmax":
        coef = attn
    elif inference == "activated":
        coef = prob * (prob > 0.2).float()
    elif inference == "argmax_test":
        coef = F.one_hot(prob.max(dim=1)[1], num_classes=agent_num).float().transpose(1, 2)
    else:
        raise ValueError("Incorrect inference mode")
    ce = coef.view(batch_size, agent_num, agent_num, 1, 1, 1)
    v = val_mat if warp_flag == 1 else val_mat.unsqueeze(2).expand(-1, -1, agent_num, -1, -1, -1)
    fuse = (ce * v).sum(1)
    fused = torch.flip(torch.cat([fuse[:, i] for i in range(agent_num)], 0), (2,))
    logits = seg_decode(fused, x1, x2, x3, sd)
    return dict(logits=logits, attn=attn, fused=fused, x4=x4) if stages else logits


# ---------------------------------------------------------------------------------------------
# Intermediate-fusion baselines (CP/models/det/base/FusionBase.py, CP/models/seg/FusionBase.py)
# ---------------------------------------------------------------------------------------------
def _pair_weight_net(x, sd, p, agent_wise=False):
    """PixelWeightedFusionSoftmax.forward (DiscoNet.py:149-155) / AgentWeightedFusion.forward
    (AgentWiseWeightedFusion.py:66-76): 1x1 convs 2C->128->32->8->1 (+ the 32x32 conv1_5), ReLU after each."""
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1_1.weight"], sd[p + "conv1_1.bias"]), sd, p + "bn1_1"))
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1_2.weight"], sd[p + "conv1_2.bias"]), sd, p + "bn1_2"))
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv1_3.weight"], sd[p + "conv1_3.bias"]), sd, p + "bn1_3"))
    x = F.relu(F.conv2d(x, sd[p + "conv1_4.weight"], sd[p + "conv1_4.bias"]))
    if agent_wise:
        x = F.relu(F.conv2d(x, sd[p + "conv1_5.weight"], sd[p + "conv1_5.bias"]))
    return x


def fuse_rule(kind, tg, nb, sd, seg=False):
    """``fusion()`` of the FusionBase subclasses on the neighbour list ``nb`` (nb[0] is the target ``tg``):
    MeanFusion.py:11-12, MaxFusion.py:20-21, SumFusion.py:20-21, CatFusion.py:22-26,
    AgentWiseWeightedFusion.py:24-43, DiscoNet.py:80-107 (det) and their seg twins."""
    if kind == "mean":
        return torch.mean(torch.stack(nb), dim=0)
    if kind == "max":
        return torch.max(torch.stack(nb), dim=0).values
    if kind == "sum":
        return torch.sum(torch.stack(nb), dim=0)
    if kind == "cat":
        p = "modulation_layer_3." if seg else "_modulation_layer_3._"
        mean = torch.mean(torch.stack(nb), dim=0)
        cat = torch.cat([tg, mean], dim=0).unsqueeze(0)
        y = F.conv2d(cat, sd[p + "conv1_1.weight"], sd[p + "conv1_1.bias"])
        return F.relu(_bn(y, sd, p + "bn1_1")).squeeze(0)
    if kind == "agent":
        w = [_pair_weight_net(torch.cat([tg, f], dim=0).unsqueeze(0), sd, "agent_weighted_fusion.", True) for f in nb]
        # torch.tensor(list of 0-dim tensors) copies the values OUT of the autograd graph (so the weight net receives no
        # gradient in the reference, AgentWiseWeightedFusion.py:24-26) and keeps their dtype
        soft = torch.squeeze(F.softmax(torch.tensor([float(t) for t in w], dtype=tg.dtype).unsqueeze(0), dim=1), 0)
        out = 0
        for k in range(len(nb)):
            out = out + soft[k] * nb[k]
        return out
    if kind == "disco":
        e = [torch.exp(torch.squeeze(_pair_weight_net(torch.cat([tg, f], dim=0).unsqueeze(0), sd,
                                                      "pixel_weighted_fusion."))) for f in nb]
        total = sum(e)
        out = 0
        for k in range(len(nb)):
            out = out + torch.div(e[k], total) * nb[k]
        return out
    raise ValueError(kind)


def fusion_stage(kind, feat_maps, trans_matrices, num_agent_tensor, sd, batch_size, agent_num, only_v2i=False,
                 seg=False):
    """The per-scene / per-agent loop of FusionBase.forward (FusionBase.py:31-63; seg/FusionBase.py:38-73) incl. the
    H flips (DetModelBase.py:71-92,53-69; SegModelBase.py:46-58,80-88).  feat_maps: [A*B,C,h,w] agent-major."""
    c, h, w = feat_maps.shape[1:]
    size = (1, c, h, w)
    feat = torch.flip(feat_maps, (2,))
    local = torch.stack([feat[batch_size * i: batch_size * (i + 1)] for i in range(agent_num)], 1)
    upd = local.clone()
    for b in range(batch_size):
        na = int(num_agent_tensor[b, 0])
        for i in range(na):
            tg = local[b, i]
            nb = [tg]
            for j in range(na):
                if j != i:
                    if only_v2i and i != 0 and j != 0:
                        continue
                    nb.append(feature_transformation(local, b, j, i, trans_matrices, size))
            upd[b, i] = fuse_rule(kind, tg, nb, sd, seg=seg)
    return torch.flip(torch.cat([upd[:, i] for i in range(agent_num)], 0), (2,))


def fusion_det_forward(kind, bevs, trans_matrices, num_agent_tensor, sd, batch_size=1, agent_num=5, only_v2i=False,
                       stages=False, layer=3):
    """det FusionBase.forward / DiscoNet.forward (kd_flag = 0 result dict)."""
    enc = encode(bevs, sd, "u_encoder.", compress_level=int("u_encoder.com_compresser.weight" in sd))
    fused = fusion_stage(kind, enc[layer], trans_matrices, num_agent_tensor, sd, batch_size, agent_num, only_v2i)
    dec_in = list(enc)
    dec_in[layer] = fused
    dec = decode(*dec_in, sd, "decoder.", kd_flag=True)
    res = heads(dec[0], sd)
    if stages:
        res = dict(res, enc=enc, fused=fused, dec=dec)
    return res


def seg_fusion_forward(kind, x, trans_matrices, num_agent_tensor, sd, agent_num=5, only_v2i=False, stages=False):
    """seg FusionBase.forward (seg/FusionBase.py:25-84), kd_flag = False."""
    x1, x2, x3, x4 = seg_encode(x, sd)
    batch_size = x.size(0) // agent_num
    fused = fusion_stage(kind, x4, trans_matrices, num_agent_tensor, sd, batch_size, agent_num, only_v2i, seg=True)
    logits = seg_decode(fused, x1, x2, x3, sd)
    return dict(logits=logits, x4=x4, fused=fused) if stages else logits


def teacher_forward(bevs, sd):
    """TeacherNet.forward == STPN_KD.forward (Backbone.py:251-257): (x_8, x_7, x_6, x_5, x_3, x_4)."""
    enc = encode(bevs, sd, "stpn.")
    dec = decode(*enc, sd, "stpn.", kd_flag=True)
    return (*dec, enc[3], enc[4])


# ---------------------------------------------------------------------------------------------
# Training step oracle: train-mode forward + vector-Jacobian product
# ---------------------------------------------------------------------------------------------
def train_step_vjp(forward, sd, upstream):
    """Run ``forward(sd)`` (a closure over one of the ``*_forward`` functions above) in training mode with every
    floating-point parameter of ``sd`` as an autograd leaf, back-propagate ``sum_k <out[k], upstream[k]>`` and return
    (outputs, {param name: grad}, sd after the step -- its BN running buffers are updated in place, as the module's are).
    What FaFModule.step needs from the model (CoDetModule.py:217-310): the loss itself stays the reference's python and
    only hands d(loss)/d(loc), d(loss)/d(cls) back."""
    work = {}
    for k, v in sd.items():
        buf = k.endswith(("running_mean", "running_var", "num_batches_tracked"))
        work[k] = v.clone() if buf else v.clone().requires_grad_(v.is_floating_point())
    with training_mode():
        out = forward(work)
    keys = sorted(upstream)
    torch.autograd.backward([out[k] for k in keys], [upstream[k] for k in keys])
    grads = {k: v.grad for k, v in work.items() if v.requires_grad and v.grad is not None}
    return {k: out[k].detach() for k in keys}, grads, {k: v.detach() for k, v in work.items()}
