"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs and weights.

The reference ships no data or checkpoints, so every parity run uses
  * occupancy BEVs   : i.i.d. Bernoulli(p) in {0,1}, agent-major ``[A*B,1,256,256,13]`` float32
                       (shape contract: V2XSimDet.py:291-302, DOC/tutorials/collaborative_models.md:25-50)
  * trans_matrices   : ``[B,A,A,4,4]`` float64, ``T[b,a,k] = P_a^-1 P_k`` for seeded SE(2) poses
                       (nuscenes_pc_util.py:230-232; the model reads ``T[b,j,i]``, DetModelBase.py:158)
  * num_agent_tensor : ``[B,A]`` int64
  * weights          : a seeded state_dict with the reference's key names/shapes
                       (Backbone.py:9-87, DetModelBase.py:268-351, convolutional_rnn/module.py:62-116);
                       conv weights are He-uniform so the signal survives ~25 layers and an
                       error anywhere is visible at the outputs; BN running stats / affine are
                       randomised so BN folding is exercised.
Everything is generated with a CPU ``torch.Generator`` so the GPU box reproduces the
exact tensors the golden fixtures were made from.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

import os
import sys

_SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "v2x-sim_b200")
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)
# the generators bench.py also uses live in the product package; re-exported here so fixtures and benchmark agree
from v2x_b200.synthetic import (BACKBONE_BNS, BACKBONE_CONVS, BOX_CODE, CATEGORY_NUM, MAP_DIMS, NUM_ANCHORS,  # noqa: E402,F401
                                _Gen, _bn, _conv, backbone_state, heads_state, make_bevs, make_poses, make_scene,
                                make_trans_matrices, plant_detections, v2vnet_det_state, fafnet_state, SEG_DOUBLE_CONVS,
                                _double_conv, seg_unet_state, seg_when2com_state, make_seg_scene, when2com_det_state)


def seg_v2vnet_state(seed=0, n_classes=8):
    """seg V2VNet (CP/models/seg/V2VNet.py:9-23): SegModelBase + Conv2dGRU(1024 -> 512)."""
    g = _Gen(seed)
    sd = seg_unet_state(seed, n_classes, g=g)
    c = 512
    bw = math.sqrt(3.0 / (2 * c * 9))
    sd["convgru.weight_ih_l0"] = g.uniform((3 * c, 2 * c, 3, 3), -bw, bw)
    sd["convgru.weight_hh_l0"] = g.uniform((3 * c, c, 3, 3), -bw, bw)
    sd["convgru.bias_ih_l0"] = g.uniform((3 * c,), -0.5, 0.5)
    sd["convgru.bias_hh_l0"] = g.uniform((3 * c,), -0.5, 0.5)
    return sd


# ---------------------------------------------------------------------------------------------
# intermediate-fusion baselines (CP/models/det/base/FusionBase.py, CP/models/seg/FusionBase.py)
# ---------------------------------------------------------------------------------------------
FUSION_KINDS = ("mean", "max", "sum", "cat", "agent", "disco")


def _pair_weight_net_state(sd, g, p, channel, agent_wise):
    """PixelWeightedFusionSoftmax (DiscoNet.py:132-147) / AgentWeightedFusion (AgentWiseWeightedFusion.py:44-64).
    The last layers are drawn positive so the ReLU'd scores are non-zero and differ between the list members (an
    all-zero score map would make every fusion weight 1/n and hide indexing errors)."""
    _conv(sd, g, p + "conv1_1", 128, 2 * channel, k=(1, 1))
    _bn(sd, g, p + "bn1_1", 128)
    _conv(sd, g, p + "conv1_2", 32, 128, k=(1, 1))
    _bn(sd, g, p + "bn1_2", 32)
    _conv(sd, g, p + "conv1_3", 8, 32, k=(1, 1))
    _bn(sd, g, p + "bn1_3", 8)
    sd[p + "conv1_4.weight"] = g.uniform((1, 8, 1, 1), 0.0, 0.8)
    sd[p + "conv1_4.bias"] = g.uniform((1,), 0.0, 0.2)
    if agent_wise:
        sd[p + "conv1_5.weight"] = g.uniform((1, 1, 32, 32), 0.0, 4.0 / 1024)
        sd[p + "conv1_5.bias"] = g.uniform((1,), 0.0, 0.2)


def _fusion_extra_state(sd, g, kind, channel, seg):
    if kind == "cat":
        p = "modulation_layer_3." if seg else "_modulation_layer_3._"
        _conv(sd, g, p + "conv1_1", channel, 2 * channel, k=(1, 1))
        _bn(sd, g, p + "bn1_1", channel)
    elif kind == "agent":
        _pair_weight_net_state(sd, g, "agent_weighted_fusion.", channel, True)
    elif kind == "disco":
        _pair_weight_net_state(sd, g, "pixel_weighted_fusion.", channel, False)


def fusion_det_state(kind, seed=0, compress_level=0, channel=256):
    """state_dict of the det FusionBase family (IntermediateModelBase.py:24-25 + the subclass' fusion net)."""
    assert kind in FUSION_KINDS
    g = _Gen(seed)
    sd = OrderedDict()
    heads_state(sd, g)
    backbone_state(sd, g, "u_encoder.", compress_level=compress_level)
    backbone_state(sd, g, "decoder.")
    _fusion_extra_state(sd, g, kind, channel, False)
    return sd


def seg_fusion_state(kind, seed=0, n_classes=8):
    """state_dict of the seg FusionBase family (SegModelBase + the subclass' fusion net at C = 512)."""
    assert kind in FUSION_KINDS
    g = _Gen(seed)
    sd = seg_unet_state(seed, n_classes, g=g)
    _fusion_extra_state(sd, g, kind, 512, True)
    return sd


def make_gt_from_detections(dets, seed=0, keep=0.7, jitter=0.25, extra=3):
    """Synthetic ground truth for the AP computation: a random ~70% of the oracle's detections, each shifted by a
    small random offset, plus a few boxes nobody predicts.  dets: list of [m,9]; returns a list of [n,8]."""
    rng = np.random.RandomState(seed)
    out = []
    for d in dets:
        d = np.asarray(d).reshape(-1, 9)
        take = rng.rand(d.shape[0]) < keep
        g = d[take, :8].reshape(-1, 4, 2) + rng.uniform(-jitter, jitter, size=(int(take.sum()), 1, 2))
        cx, cy = rng.uniform(-28, 28, size=(2, extra))
        base = np.array([[-1.0, 2.0], [1.0, 2.0], [1.0, -2.0], [-1.0, -2.0]])
        ex = base[None] + np.stack([cx, cy], axis=-1)[:, None, :]
        out.append(np.concatenate([g, ex], axis=0).reshape(-1, 8))
    return out


# ---------------------------------------------------------------------------------------------
# sparse voxel inputs (CP/datasets/V2XSimDet.py:291-302)
# ---------------------------------------------------------------------------------------------
def make_voxel_indices(num_maps, seed=0, points=3000, dims=(256, 256, 13), duplicates=True):
    """Per-map int32 [n, 3] voxel index lists like the dataset's ``voxel_indices_0`` (duplicates allowed, as a LiDAR
    sweep quantised to voxels produces them)."""
    rng = np.random.RandomState(seed)
    out = []
    for m in range(num_maps):
        n = points + 97 * m
        idx = np.stack([rng.randint(0, d, size=n) for d in dims], 1).astype(np.int32)
        if duplicates:
            idx = np.concatenate([idx, idx[: n // 10]], 0)
        out.append(idx)
    return out


def densify_voxels(indices, dims=(256, 256, 13)):
    """Restatement of V2XSimDet.py:294-302 for one map: zeros(bool) -> scatter 1 -> np.rot90(., 3) -> float32."""
    curr_voxels = np.zeros(dims, dtype=bool)
    curr_voxels[indices[:, 0], indices[:, 1], indices[:, 2]] = 1
    curr_voxels = np.rot90(curr_voxels, 3)
    return np.ascontiguousarray(curr_voxels).astype(np.float32)


def voxel_rows(index_lists):
    """[(n_m, 3)] per map -> one int32 [sum n_m, 4] tensor of (map, i0, i1, i2) rows for v2x_voxelize_fwd."""
    rows = [np.concatenate([np.full((len(ix), 1), m, dtype=np.int32), ix], 1) for m, ix in enumerate(index_lists)]
    return torch.from_numpy(np.concatenate(rows, 0))
