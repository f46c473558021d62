"""Known-answer tests of the post-processing oracle (oracle/postproc.py): the reference's NMS / AP code
(CP/utils/postprocess.py, mean_ap.py) cannot import here (shapely, mmcv), so the restatement is pinned by
closed-form cases instead."""
import math

import numpy as np
import torch

from oracle import postproc as pp


def sq(cx, cy, s, ang=0.0):
    c = np.array([[-s, s], [s, s], [s, -s], [-s, -s]], dtype=np.float64) / 2
    r = np.array([[math.cos(ang), -math.sin(ang)], [math.sin(ang), math.cos(ang)]])
    return c @ r.T + np.array([cx, cy])


def test_quad_iou_closed_forms():
    a = sq(0, 0, 2)
    # identical, disjoint, half-shifted, contained
    assert abs(pp.quad_iou(a, a[None])[0] - 1.0) < 1e-6
    assert pp.quad_iou(a, sq(5, 0, 2)[None])[0] == 0.0
    assert abs(pp.quad_iou(a, sq(1, 0, 2)[None])[0] - (2.0 / 6.0)) < 1e-6
    assert abs(pp.quad_iou(a, sq(0, 0, 1)[None])[0] - 0.25) < 1e-6
    # unit-area squares rotated by 45 degrees about a common centre intersect in a regular octagon of area
    # 2(sqrt2 - 1); union = 2 - that
    b = sq(0, 0, 1, math.pi / 4)
    inter = 2 * (math.sqrt(2) - 1)
    assert abs(pp.quad_iou(sq(0, 0, 1), b[None])[0] - inter / (2 - inter)) < 1e-6
    # orientation (cw / ccw vertex order) must not matter; touching edges have zero overlap
    assert abs(pp.quad_iou(a[::-1], sq(1, 0, 2)[None])[0] - (2.0 / 6.0)) < 1e-6
    assert pp.quad_iou(a, sq(2, 0, 2)[None])[0] < 1e-9


def test_quad_iou_matches_rasterised_area():
    rng = np.random.RandomState(0)
    for _ in range(20):
        a = sq(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1, 3), rng.uniform(0, math.pi))
        b = sq(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1, 3), rng.uniform(0, math.pi))
        # Monte-Carlo-free check: fine grid rasterisation
        g = np.linspace(-4, 4, 801)
        X, Y = np.meshgrid(g, g)
        P = np.stack([X.ravel(), Y.ravel()], -1)

        def inside(q):
            q = pp._ccw(q[None])[0]
            m = np.ones(len(P), bool)
            for i in range(4):
                p0, p1 = q[i], q[(i + 1) % 4]
                m &= ((p1[0] - p0[0]) * (P[:, 1] - p0[1]) - (p1[1] - p0[1]) * (P[:, 0] - p0[0])) >= 0
            return m
        ia, ib = inside(a), inside(b)
        ref = (ia & ib).sum() / max((ia | ib).sum(), 1)
        assert abs(pp.quad_iou(a, b[None])[0] - ref) < 2e-2


def test_corners_and_decode():
    anchors = pp.init_anchors()
    assert anchors.shape == (256, 256, 6, 6)
    # obj_util.py:611-633: centre of cell (i=0, j=0) is (-31.875, -31.875); x follows the W index
    assert np.allclose(anchors[0, 0, 0, :2], [-31.875, -31.875]) and np.allclose(anchors[3, 7, 2, :2], [-30.125, -31.125])
    # zero encoding decodes to the anchor itself (detection_util.py:385-398)
    a = anchors.reshape(-1, 6)[[0, 1, 2]]
    dec = pp.decode_boxes(np.zeros((3, 6), np.float32) + np.array([0, 0, 0, 0, 0, 1], np.float32), a)
    assert np.allclose(dec, a, atol=1e-6)
    # un-rotated 2 x 4 box: corner order x0y1, x1y1, x1y0, x0y0 (obj_util.py:300-316)
    c = pp.corners_of(np.zeros((1, 2)), np.array([[2.0, 4.0]]), np.array([[0.0, 1.0]]))
    assert np.allclose(c[0], [[-1, 2], [1, 2], [1, -2], [-1, -2]])
    # sin=1, cos=0 rotates "clockwise when the angle is positive" (obj_util.py:344-359): (x,y) -> (y,-x)
    c = pp.corners_of(np.zeros((1, 2)), np.array([[2.0, 4.0]]), np.array([[1.0, 0.0]]))
    assert np.allclose(c[0], [[2, 1], [2, -1], [-2, -1], [-2, 1]])


def test_nms_order_and_threshold():
    boxes = np.stack([sq(0, 0, 2), sq(0.5, 0, 2), sq(10, 0, 2), sq(10.2, 0, 2), sq(20, 0, 2)]).astype(np.float32)
    scores = np.array([0.9, 0.95, 0.8, 0.75, 0.6], dtype=np.float32)
    # 0.6 is filtered (<= 0.7); the higher-scored box of each overlapping pair wins; output in score order
    assert pp.non_max_suppression(boxes, scores, 0.01).tolist() == [1, 2]
    # a looser threshold keeps boxes whose IoU does not exceed it
    assert pp.non_max_suppression(boxes, scores, 0.9).tolist() == [1, 0, 2, 3]


def test_average_precision_hand_cases():
    # perfect detector
    assert abs(pp.average_precision(np.array([0.5, 1.0]), np.array([1.0, 1.0])) - 1.0) < 1e-9
    # tp, fp, tp with 2 gts: recalls .5 .5 1, precisions 1 .5 2/3 -> 0.5*1 + 0.5*(2/3)
    ap = pp.average_precision(np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2.0 / 3.0]))
    assert abs(ap - (0.5 + 0.5 * 2.0 / 3.0)) < 1e-9


def test_eval_map_matching():
    gt = [np.stack([sq(0, 0, 2), sq(10, 0, 2)]).reshape(-1, 8)]
    det = [np.concatenate([np.stack([sq(0.1, 0, 2), sq(0.2, 0, 2), sq(30, 0, 2), sq(10, 0.1, 2)]).reshape(-1, 8),
                           np.array([[0.9], [0.85], [0.8], [0.75]])], axis=1)]
    # order by score: tp, fp (duplicate of a covered gt), fp (no gt), tp
    ap, info = pp.eval_map(det, gt, 0.5)
    assert info["num_gts"] == 2 and info["num_dets"] == 4
    assert np.allclose(info["recall"], [0.5, 0.5, 0.5, 1.0]) and np.allclose(info["precision"], [1, 0.5, 1 / 3, 0.5])
    assert abs(ap - (0.5 * 1.0 + 0.5 * 0.5)) < 1e-6
    # no detections at all -> AP 0; no gts in an image -> every detection there is a false positive
    assert pp.eval_map([np.zeros((0, 9))], gt, 0.5)[0] == 0.0
    ap2, info2 = pp.eval_map(det + det, gt + [np.zeros((0, 8))], 0.5)
    assert info2["num_dets"] == 8 and ap2 < ap


def test_planted_pipeline_on_oracle_outputs():
    """softmax -> decode -> NMS -> AP on planted-head logits: non-empty, deterministic, AP in (0, 1]."""
    from oracle import synth
    g = torch.Generator().manual_seed(0)
    n = 2
    cls = torch.randn((n, 256 * 256 * 6, 2), generator=g)
    loc = torch.randn((n, 256, 256, 6, 1, 6), generator=g) * 0.05
    loc[..., 5] += 1.0
    sd = {"classification.conv2.bias": torch.zeros(12), "regression.box_prediction.3.weight": torch.ones(36, 32, 1, 1),
          "regression.box_prediction.3.bias": torch.ones(36)}
    sd2 = synth.plant_detections(sd, cls, per_agent=100)
    shift = sd2["classification.conv2.bias"][1].item()
    assert sd2["classification.conv2.bias"][0].item() == 0.0 and abs(sd2["regression.box_prediction.3.bias"][0].item() - 0.05) < 1e-7
    assert sd2["regression.box_prediction.3.bias"][5].item() == 1.0
    cls2 = cls.clone()
    cls2[..., 1] += shift
    dets, sel = pp.detections_of(loc.numpy(), cls2.numpy())
    n_cand = int((pp.softmax_fg(cls2.numpy())[..., 0] > 0.7).sum())
    assert 190 <= n_cand <= 210
    assert all(0 < len(s) <= 110 for s in sel)
    gts = synth.make_gt_from_detections(dets, seed=1)
    ap5, _ = pp.eval_map(dets, gts, 0.5)
    ap7, _ = pp.eval_map(dets, gts, 0.7)
    assert 0.3 < ap7 <= ap5 <= 1.0


def _clip_iou_port(qa, qb):
    """Line-by-line python port of quad_iou_f64 (csrc/postproc_kernels.cu): Sutherland-Hodgman clip of A by the four
    half-planes of B (B made counter-clockwise), shoelace areas.  Kept in step with the CUDA source so that the
    ALGORITHM the GPU NMS runs is checked on CPU against the oracle's independent construction (vertices-inside +
    edge crossings + angular sort)."""
    a = [[float(qa[i][0]), float(qa[i][1])] for i in range(4)]
    b = [[float(qb[i][0]), float(qb[i][1])] for i in range(4)]

    def signed_area2(p):
        n = len(p)
        return sum(p[i][0] * p[(i + 1) % n][1] - p[(i + 1) % n][0] * p[i][1] for i in range(n))

    sb = signed_area2(b)
    area_a, area_b = 0.5 * abs(signed_area2(a)), 0.5 * abs(sb)
    if sb < 0.0:
        b[1], b[3] = b[3], b[1]
    n = 4
    for e in range(4):
        if n == 0:
            break
        ex0, ey0 = b[e]
        dx, dy = b[(e + 1) & 3][0] - ex0, b[(e + 1) & 3][1] - ey0
        tmp = []
        for i in range(n):
            j = 0 if i + 1 == n else i + 1
            ci = dx * (a[i][1] - ey0) - dy * (a[i][0] - ex0)
            cj = dx * (a[j][1] - ey0) - dy * (a[j][0] - ex0)
            if ci >= 0.0 and len(tmp) < 8:
                tmp.append([a[i][0], a[i][1]])
            if (ci >= 0.0) != (cj >= 0.0) and len(tmp) < 8:
                t = ci / (ci - cj)
                tmp.append([a[i][0] + t * (a[j][0] - a[i][0]), a[i][1] + t * (a[j][1] - a[i][1])])
        a, n = tmp, len(tmp)
    if n < 3:
        return 0.0
    inter = 0.5 * abs(signed_area2(a))
    uni = area_a + area_b - inter
    return inter / uni if uni > 0.0 else 0.0


def test_gpu_iou_algorithm_port_matches_oracle_on_random_quads():
    rng = np.random.RandomState(7)
    worst = 0.0
    for k in range(1500):
        a = sq(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(0.5, 4), rng.uniform(0, math.pi))
        b = sq(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(0.5, 4), rng.uniform(0, math.pi))
        a[:, 0] *= rng.uniform(0.5, 2.0)            # rectangles, not only squares
        if k % 3 == 0:
            b = b[::-1].copy()                       # clockwise vertex order
        if k % 50 == 0:
            b = a.copy()                             # identical boxes
        if k % 70 == 0:
            b = a + np.array([a[1, 0] - a[0, 0], a[1, 1] - a[0, 1]])   # sharing one edge exactly
        ref = float(pp.quad_iou(a, b[None])[0])
        got = _clip_iou_port(a.astype(np.float32), b.astype(np.float32))
        worst = max(worst, abs(got - ref))
    assert worst < 5e-6, worst          # float32 corners on the port's side (as the kernel receives them)
