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


# ---------------------------------------------------------------------------------------------------------------------
# Exact-arithmetic pin of the polygon IoU.  The reference computes IoU with shapely
# (postprocess.py:42-52: Polygon(a).intersection(Polygon(b)).area / union area); shapely / GEOS are absent from this
# image, so the strongest pin available is an INDEPENDENT implementation in exact rational arithmetic: Sutherland-Hodgman
# clipping over fractions.Fraction of the float64 corner values (every float is a rational, so the clip vertices, the
# shoelace areas and the IoU are exact -- no tolerance, no epsilon).  oracle/postproc.quad_iou (float64, a different
# construction: vertices-inside + edge crossings + angular sort) must agree with it to 1e-12.
# ---------------------------------------------------------------------------------------------------------------------
def _exact_iou(qa, qb):
    from fractions import Fraction as Fr

    def F(p):
        return [(Fr(float(x)), Fr(float(y))) for x, y in p]

    def area2(p):   # twice the signed area
        n = len(p)
        return sum(p[i][0] * p[(i + 1) % n][1] - p[(i + 1) % n][0] * p[i][1] for i in range(n))

    a, b = F(qa), F(qb)
    if area2(a) < 0:
        a = a[::-1]
    if area2(b) < 0:
        b = b[::-1]
    poly = a
    for e in range(4):
        if not poly:
            break
        (x0, y0), (x1, y1) = b[e], b[(e + 1) % 4]
        dx, dy = x1 - x0, y1 - y0
        side = [dx * (p[1] - y0) - dy * (p[0] - x0) for p in poly]
        out = []
        for i in range(len(poly)):
            j = (i + 1) % len(poly)
            if side[i] >= 0:
                out.append(poly[i])
            if (side[i] > 0 and side[j] < 0) or (side[i] < 0 and side[j] > 0):
                t = side[i] / (side[i] - side[j])
                out.append((poly[i][0] + t * (poly[j][0] - poly[i][0]), poly[i][1] + t * (poly[j][1] - poly[i][1])))
        poly = out
    inter = abs(area2(poly)) / 2 if len(poly) >= 3 else Fr(0)
    union = abs(area2(a)) / 2 + abs(area2(b)) / 2 - inter
    return inter / union if union > 0 else Fr(0)


def _random_quad(rng):
    q = sq(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(0.5, 4), rng.uniform(0, math.pi))
    q[:, 0] *= rng.uniform(0.5, 2.0)
    return q


def test_quad_iou_matches_exact_rational_clipping_on_random_quads():
    rng = np.random.RandomState(11)
    worst = 0.0
    for k in range(1600):
        a, b = _random_quad(rng), _random_quad(rng)
        if k % 3 == 0:
            b = b[::-1].copy()                                            # clockwise vertex order
        if k % 50 == 0:
            b = a.copy()                                                  # identical boxes -> exactly 1
        if k % 70 == 0:
            b = a + np.array([a[1, 0] - a[0, 0], a[1, 1] - a[0, 1]])      # sharing one edge -> exactly 0
        if k % 90 == 0:
            b = a * 0.5 + a.mean(0) * 0.5                                 # strictly nested
        exact = _exact_iou(a, b)
        got = float(pp.quad_intersection_area(a[None], b[None])[0])
        union = float(pp._area(a[None])[0] + pp._area(b[None])[0]) - got
        worst = max(worst, abs(got / union - float(exact)))
        # the float32 value the NMS compares (postprocess.py:41: iou array is float32)
        assert abs(float(pp.quad_iou(a, b[None])[0]) - float(np.float32(float(exact)))) <= 2e-7
    assert worst < 1e-12, worst


def test_quad_iou_around_the_nms_threshold():
    """Adversarial cases at the 0.01 IoU threshold of apply_nms_det (detection_util.py:357-359 -> postprocess.py:106
    ``iou > threshold``): two 2 x 4 boxes slid apart until their IoU is a hair above / below 0.01; the suppress
    decision of the restated NMS must equal the decision made on the exact rational IoU (compared in float32, as the
    reference's float32 iou array is)."""
    thr = np.float32(0.01)
    base = np.array([[-1.0, -2.0], [-1.0, 2.0], [1.0, 2.0], [1.0, -2.0]])
    w, h = 2.0, 4.0
    decided = 0
    for ang in (0.0, 0.3, 1.1, math.pi / 4):
        c, s = math.cos(ang), math.sin(ang)
        rot = np.array([[c, -s], [s, c]])
        a = base @ rot.T
        # axis-aligned overlap of width d: IoU = d*h / (2*w*h - d*h) = thr  ->  d = 2*w*thr / (1 + thr)
        d0 = 2 * w * 0.01 / 1.01
        for rel in (-1e-3, -1e-5, -1e-7, -1e-9, 0.0, 1e-9, 1e-7, 1e-5, 1e-3):
            d = d0 * (1 + rel)
            b = (base + np.array([w - d, 0.0])) @ rot.T
            exact = _exact_iou(a, b)
            want = np.float32(float(exact)) > thr
            got32 = pp.quad_iou(a, b[None])[0]
            assert abs(float(got32) - float(exact)) < 1e-9
            # the float64 IoU is within 1e-12 of the exact one, so both round to the same float32 (unless the exact value
            # sat within 1e-12 of a float32 rounding boundary, which none of these do): the decisions must agree
            assert got32 == np.float32(float(exact)), (ang, rel, float(exact), float(got32))
            assert bool(got32 > thr) == bool(want)
            boxes = np.stack([a, b]).astype(np.float64)
            keep = pp.non_max_suppression(boxes, np.array([0.9, 0.8], dtype=np.float32), 0.01)
            assert (len(keep) == 1) == bool(want)
            decided += int(bool(want))
    assert 12 <= decided <= 24     # both sides of the threshold were exercised
