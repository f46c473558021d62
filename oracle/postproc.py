"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's detection post-processing and mAP.

The reference evaluates a detector by  softmax -> box decode -> corners -> score filter + polygon NMS ->
per-class TP/FP matching -> area-mode AP  (test_codet.py:295-466).  None of those modules import in this
image (postprocess.py:9 needs shapely, mean_ap.py:3-4 need mmcv / terminaltables), so they are restated here
in numpy, function by function:

  init_anchors          CP/utils/obj_util.py:611-633 (init_anchors_no_check), sizes Config.py:154-163
  decode_boxes          CP/utils/detection_util.py:376-398 (bev_box_decode_torch)
  corners_of            CP/utils/obj_util.py:270-341 (corners_nd, center_to_corner_box2d), :344-359 (rotation_2d)
  quad_iou              shapely ``Polygon.intersection / union`` of two convex quads (postprocess.py:41-53);
                        here: exact convex-polygon clipping (vertices-inside + edge crossings, shoelace)
  non_max_suppression   CP/utils/postprocess.py:72-113 (score > 0.7, descending, IoU > threshold suppressed)
  apply_nms_det         CP/utils/detection_util.py:256-373 (softmax[...,1:], decode, corners, NMS with 0.01)
  tpfp / eval_map       CP/utils/mean_ap.py:51-178, :181-304, average_precision :8-49 (mode="area")

PARITY: shapely / mmcv themselves are absent from this image.  The polygon IoU -- the one piece of third-party
arithmetic (shapely ``Polygon.intersection``, postprocess.py:42-52) -- is pinned instead by an independent
EXACT-RATIONAL implementation (fractions.Fraction Sutherland-Hodgman, tests/test_postproc.py::_exact_iou): 1600 random
quads (cw / ccw, identical, edge-sharing, nested) agree to 1e-12, and adversarial pairs a hair above / below the 0.01
NMS threshold give the same float32 value and the same suppress decision.  That is the strongest pin available without
the dependency; the rest is pinned by analytic known-answer cases (closed-form overlaps, VOC AP hand examples).
"""
from __future__ import annotations

import math

import numpy as np

AREA_EXTENTS = ((-32.0, 32.0), (-32.0, 32.0), (-8.0, -3.0))   # Config.py:78-84
VOXEL_SIZE = (0.25, 0.25, 0.4)                                  # Config.py:76
ANCHOR_SIZE = np.asarray([[2.0, 4.0, 0.0], [2.0, 4.0, math.pi / 2.0], [2.0, 4.0, -math.pi / 4.0],
                          [3.0, 12.0, 0.0], [3.0, 12.0, math.pi / 2.0], [3.0, 12.0, -math.pi / 4.0]])  # Config.py:154-163
SCORE_FILTER = 0.7   # postprocess.py:84
NMS_IOU = 0.01       # detection_util.py:357-359


def init_anchors(box_code_size=6, anchor_size=ANCHOR_SIZE):
    """[H, W, anchors, 6] = (x, y, w, h, sin, cos); obj_util.py:611-633."""
    w_range = math.ceil((AREA_EXTENTS[0][1] - AREA_EXTENTS[0][0]) / VOXEL_SIZE[0])
    h_range = math.ceil((AREA_EXTENTS[1][1] - AREA_EXTENTS[1][0]) / VOXEL_SIZE[1])
    a = np.zeros((h_range, w_range, len(anchor_size), box_code_size))
    a[:, :, :, 2:4] = anchor_size[:, :2]
    a[:, :, :, 4] = np.sin(anchor_size[:, 2])
    a[:, :, :, 5] = np.cos(anchor_size[:, 2])
    jj = np.arange(w_range) * VOXEL_SIZE[0] + AREA_EXTENTS[0][0] + VOXEL_SIZE[0] / 2.0
    ii = np.arange(h_range) * VOXEL_SIZE[0] + AREA_EXTENTS[0][0] + VOXEL_SIZE[1] / 2.0   # (sic) x extent, obj_util.py:627
    a[:, :, :, 0] = jj[None, :, None]
    a[:, :, :, 1] = ii[:, None, None]
    return a


def decode_boxes(enc, anchors):
    """detection_util.py:376-398; enc, anchors [N,6] -> [N,6] (x, y, w, h, sin, cos), float32 like the reference."""
    enc = np.asarray(enc, dtype=np.float32)
    anchors = np.asarray(anchors, dtype=np.float32)
    xa, ya, wa, ha, sina, cosa = [anchors[:, i] for i in range(6)]
    xp, yp, wp, hp, sinp, cosp = [enc[:, i] for i in range(6)]
    h = ha / np.exp(hp)
    w = wa / np.exp(wp)
    x = xa - w * xp
    y = ya - h * yp
    sin = sina * cosp + cosa * sinp
    cos = cosa * cosp - sina * sinp
    return np.stack([x, y, w, h, sin, cos], axis=-1)


def corners_of(centers, dims, angles):
    """obj_util.py:270-359: [N,2],[N,2],[N,2](sin,cos) -> [N,4,2], corner order x0y1, x1y1, x1y0, x0y0 of the
    un-rotated box, then rotated by [[cos,-sin],[sin,cos]]^T and shifted."""
    norm = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=dims.dtype) - 0.5     # corners_norm[[0,1,3,2]] - origin
    c = dims.reshape(-1, 1, 2) * norm.reshape(1, 4, 2)
    c = c[:, [1, 2, 3, 0], :]
    s, co = angles[:, 0], angles[:, 1]
    rot_t = np.stack([np.stack([co, -s]), np.stack([s, co])])          # [2,2,N]
    c = np.einsum("aij,jka->aik", c, rot_t)
    return c + centers.reshape(-1, 1, 2)


def _area(poly):
    x, y = poly[..., 0], poly[..., 1]
    return 0.5 * np.abs(np.sum(x * np.roll(y, -1, axis=-1) - y * np.roll(x, -1, axis=-1), axis=-1))


def _ccw(q):
    x, y = q[..., 0], q[..., 1]
    signed = np.sum(x * np.roll(y, -1, axis=-1) - y * np.roll(x, -1, axis=-1), axis=-1)
    out = q.copy()
    flip = signed < 0
    out[flip] = q[flip][:, ::-1]
    return out


def quad_intersection_area(a, b):
    """Intersection area of convex quads a[P,4,2] and b[P,4,2] (pairwise), float64.
    Vertex set of the intersection = A's vertices inside B, B's inside A, and all edge crossings; it is convex,
    so sorting the points by angle about their centroid and applying the shoelace formula is exact."""
    a = _ccw(np.asarray(a, dtype=np.float64))
    b = _ccw(np.asarray(b, dtype=np.float64))
    P = a.shape[0]
    if P == 0:
        return np.zeros((0,))
    eps = 1e-12

    def inside(pts, poly):  # pts [P,4,2] inside convex ccw poly [P,4,2] (boundary counts)
        p0 = poly[:, None, :, :]
        p1 = np.roll(poly, -1, axis=1)[:, None, :, :]
        q = pts[:, :, None, :]
        cross = (p1[..., 0] - p0[..., 0]) * (q[..., 1] - p0[..., 1]) - (p1[..., 1] - p0[..., 1]) * (q[..., 0] - p0[..., 0])
        return np.all(cross >= -eps, axis=-1)

    a_in, b_in = inside(a, b), inside(b, a)
    # edge crossings: a_i + t (a_{i+1} - a_i) = b_j + u (b_{j+1} - b_j)
    a0 = a[:, :, None, :]
    da = (np.roll(a, -1, axis=1) - a)[:, :, None, :]
    b0 = b[:, None, :, :]
    db = (np.roll(b, -1, axis=1) - b)[:, None, :, :]
    den = da[..., 0] * db[..., 1] - da[..., 1] * db[..., 0]
    diff = b0 - a0
    ok = np.abs(den) > eps
    den_s = np.where(ok, den, 1.0)
    t = (diff[..., 0] * db[..., 1] - diff[..., 1] * db[..., 0]) / den_s
    u = (diff[..., 0] * da[..., 1] - diff[..., 1] * da[..., 0]) / den_s
    hit = ok & (t >= 0) & (t <= 1) & (u >= 0) & (u <= 1)
    xpts = a0 + t[..., None] * da                                              # [P,4,4,2]
    pts = np.concatenate([a, b, xpts.reshape(P, 16, 2)], axis=1)               # [P,24,2]
    valid = np.concatenate([a_in, b_in, hit.reshape(P, 16)], axis=1)           # [P,24]
    cnt = valid.sum(axis=1)
    w = valid[..., None].astype(np.float64)
    cen = (pts * w).sum(axis=1) / np.maximum(cnt, 1)[:, None]
    ang = np.arctan2(pts[..., 1] - cen[:, None, 1], pts[..., 0] - cen[:, None, 0])
    ang = np.where(valid, ang, np.inf)                                          # invalid points sort last
    order = np.argsort(ang, axis=1)
    pts_s = np.take_along_axis(pts, order[..., None], axis=1)
    valid_s = np.take_along_axis(valid, order, axis=1)
    first = pts_s[:, :1, :]
    pts_s = np.where(valid_s[..., None], pts_s, first)                          # pad with the first vertex: zero-area terms
    area = _area(pts_s)
    return np.where(cnt >= 3, area, 0.0)


def quad_iou(box, boxes):
    """postprocess.py:41-53 compute_iou: IoU of one quad [4,2] with quads [K,4,2] -> float32 [K]."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4, 2)
    if boxes.shape[0] == 0:
        return np.zeros((0,), dtype=np.float32)
    a = np.broadcast_to(np.asarray(box, dtype=np.float64).reshape(1, 4, 2), boxes.shape)
    inter = quad_intersection_area(a, boxes)
    union = _area(a) + _area(boxes) - inter
    return (inter / np.maximum(union, 1e-300)).astype(np.float32)


def non_max_suppression(boxes, scores, threshold, score_filter=SCORE_FILTER):
    """postprocess.py:72-113: keep scores > 0.7, visit in descending score order, drop IoU > threshold."""
    assert boxes.shape[0] > 0
    boxes = boxes.astype(np.float32) if boxes.dtype.kind != "f" else boxes
    fil_id = np.where(scores > score_filter)[0]
    ixs_sort = scores[fil_id].argsort()[::-1]
    ixs = fil_id[ixs_sort]
    polys = boxes[ixs]
    it = np.arange(len(ixs))
    pick = []
    while len(it) > 0:
        i = it[0]
        pick.append(ixs[i])
        iou = quad_iou(polys[i], polys[it[1:]])
        remove = np.where(iou > threshold)[0] + 1
        it = np.delete(it, remove)
        it = np.delete(it, 0)
    return np.array(pick, dtype=np.int32)


def softmax_fg(cls):
    """F.softmax(cls_preds, -1)[..., 1:] (detection_util.py:275) for [..., 2] logits -> [..., 1] float32."""
    cls = np.asarray(cls, dtype=np.float32)
    m = cls.max(axis=-1, keepdims=True)
    e = np.exp(cls - m)
    return (e / e.sum(axis=-1, keepdims=True))[..., 1:]


def apply_nms_det(loc, cls, anchors=None):
    """detection_util.py:256-373 for one agent: loc [H,W,A,1,6], cls [H*W*A,2] ->
    dict(pred [K,4,2] corners, score [K], selected_idx [K]) (class 1 of the binary config, only_det)."""
    if anchors is None:
        anchors = init_anchors()
    scores = softmax_fg(cls)[:, 0]
    enc = np.asarray(loc, dtype=np.float32).reshape(-1, 6)
    cand = np.where(scores > SCORE_FILTER)[0]   # only these can survive NMS (postprocess.py:84): decode just them
    corners = np.zeros((enc.shape[0], 4, 2), dtype=np.float32)
    if len(cand):
        dec = decode_boxes(enc[cand], anchors.reshape(-1, 6)[cand])
        corners[cand] = corners_of(dec[:, :2], dec[:, 2:4], dec[:, 4:])
    sel = non_max_suppression(corners, scores, NMS_IOU) if len(cand) else np.zeros((0,), dtype=np.int32)
    return {"pred": corners[sel], "score": scores[sel], "selected_idx": sel}


# ---------------------------------------------------------------------------------------------------------
# mean_ap.py
# ---------------------------------------------------------------------------------------------------------
def average_precision(recalls, precisions):
    """mean_ap.py:8-49, mode='area', single scale."""
    mrec = np.concatenate([[0.0], recalls, [1.0]])
    mpre = np.concatenate([[0.0], precisions, [0.0]])
    for i in range(mpre.shape[0] - 1, 0, -1):
        mpre[i - 1] = max(mpre[i - 1], mpre[i])
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def tpfp(det, gt, iou_thr):
    """mean_ap.py:51-178 (no ignore boxes, no area ranges): det [m,9] = 8 corner coords + score, gt [n,8]."""
    m, n = det.shape[0], gt.shape[0]
    tp, fp = np.zeros(m, dtype=np.float32), np.zeros(m, dtype=np.float32)
    if n == 0:
        fp[:] = 1
        return tp, fp
    if m == 0:
        return tp, fp
    gtc = gt[:, :8].reshape(n, 4, 2).astype(np.float32)
    dc = det[:, :8].reshape(m, 4, 2).astype(np.float32)
    ious = np.stack([quad_iou(g, dc) for g in gtc], axis=0).T       # [m, n]
    ious_max, ious_arg = ious.max(axis=1), ious.argmax(axis=1)
    covered = np.zeros(n, dtype=bool)
    for i in np.argsort(-det[:, -1]):
        if ious_max[i] >= iou_thr:
            g = ious_arg[i]
            if not covered[g]:
                covered[g] = True
                tp[i] = 1
            else:
                fp[i] = 1
        else:
            fp[i] = 1
    return tp, fp


def eval_map(det_results, annotations, iou_thr=0.5):
    """mean_ap.py:181-304 for the single positive class of the binary config.
    det_results: list (images) of [m,9]; annotations: list of [n,8].  Returns (mAP, dict)."""
    assert len(det_results) == len(annotations)
    tps, fps, num_gts = [], [], 0
    for d, g in zip(det_results, annotations):
        t, f = tpfp(np.asarray(d, dtype=np.float64).reshape(-1, 9), np.asarray(g, dtype=np.float64).reshape(-1, 8), iou_thr)
        tps.append(t)
        fps.append(f)
        num_gts += np.asarray(g).reshape(-1, 8).shape[0]
    dets = np.vstack([np.asarray(d, dtype=np.float64).reshape(-1, 9) for d in det_results])
    order = np.argsort(-dets[:, -1])
    tp = np.cumsum(np.hstack(tps)[order])
    fp = np.cumsum(np.hstack(fps)[order])
    eps = np.finfo(np.float32).eps
    recalls = tp / max(num_gts, eps)
    precisions = tp / np.maximum(tp + fp, eps)
    ap = average_precision(recalls, precisions) if num_gts > 0 else 0.0
    return ap, {"num_gts": num_gts, "num_dets": dets.shape[0], "recall": recalls, "precision": precisions, "ap": ap}


def detections_of(loc, cls, anchors=None):
    """Per-agent [m,9] detection rows (8 corner coordinates + score) as test_codet.py:300-321 assembles them."""
    n = loc.shape[0]
    anchors = init_anchors() if anchors is None else anchors
    out, sel = [], []
    for a in range(n):
        r = apply_nms_det(loc[a], cls[a], anchors)
        out.append(np.concatenate([r["pred"].reshape(-1, 8), r["score"].reshape(-1, 1)], axis=1))
        sel.append(r["selected_idx"])
    return out, sel
