"""CPU proofs (against torch) of the host-side weight transforms in v2x_b200/transforms.py: the parity decomposition of
a 3x3 conv over a nearest-2x-upsampled map, and the data gradients of the path's stride-1 / stride-2 3x3 convs as
stride-1 correlations.  float64, so the identities hold to round-off."""
import pytest
import torch
import torch.nn.functional as F

from v2x_b200 import transforms as T


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


@pytest.mark.parametrize("n,ci,co,h,w", [(1, 3, 4, 5, 6), (2, 8, 5, 4, 4), (1, 1, 1, 1, 1), (1, 4, 4, 7, 2)])
def test_upsample_conv_parity_decomposition(n, ci, co, h, w):
    """conv3x3(pad 1) of F.interpolate(x, 2) (Backbone.py:173-178) == four 2x2 correlations of x, interleaved."""
    x, wt = _rand((n, ci, h, w), 1), _rand((co, ci, 3, 3), 2)
    ref = F.conv2d(F.interpolate(x, scale_factor=(2, 2)), wt, padding=1)
    got = T.upsample_conv_parity_apply(x, wt)
    assert got.shape == ref.shape and (got - ref).abs().max().item() < 1e-12
    weff = T.upsample_conv_parity_weights(wt)
    assert set(weff) == {(0, 0), (0, 1), (1, 0), (1, 1)} and all(v.shape == (co, ci, 2, 2) for v in weff.values())
    # every class sums all nine taps: a constant input far from the border gives the same value on all four parities
    total = wt.sum((2, 3))
    for v in weff.values():
        assert (v.sum((2, 3)) - total).abs().max().item() < 1e-12


@pytest.mark.parametrize("k", [1, 3])
def test_dgrad_stride1_is_a_correlation_with_rotated_transposed_weights(k):
    x = _rand((2, 5, 6, 7), 3).requires_grad_(True)
    wt = _rand((4, 5, k, k), 4)
    y = F.conv2d(x, wt, padding=k // 2)
    dy = _rand(tuple(y.shape), 5)
    (ref,) = torch.autograd.grad(y, x, dy)
    got = F.conv2d(dy, T.dgrad_weights_stride1(wt), padding=k // 2)
    assert (got - ref).abs().max().item() < 1e-12


@pytest.mark.parametrize("h,w", [(4, 6), (8, 8), (2, 2)])
def test_dgrad_stride2_parity_decomposition(h, w):
    """Data gradient of the path's 3x3 / stride 2 / pad 1 convs (Backbone.py:25,28,31,34) through four stride-1
    sub-correlations of dy, one per input parity class: 1 + 2 + 2 + 4 = 9 taps."""
    x = _rand((2, 3, h, w), 6).requires_grad_(True)
    wt = _rand((5, 3, 3, 3), 7)
    y = F.conv2d(x, wt, stride=2, padding=1)
    dy = _rand(tuple(y.shape), 8)
    (ref,) = torch.autograd.grad(y, x, dy)
    got = T.dgrad_stride2_apply(dy, wt, (h, w))
    assert (got - ref).abs().max().item() < 1e-12
    parts = T.dgrad_parity_weights_stride2(wt)
    assert sum(v[0].shape[2] * v[0].shape[3] for v in parts.values()) == 9


def test_gru_zero_hidden_gate_backward_matches_autograd():
    a = [_rand((2, 6, 4, 5), s).requires_grad_(True) for s in (10, 11, 12)]
    b_hn = _rand((6,), 13).requires_grad_(True)
    h, r, z, n = T.gru_zero_hidden_gates(a[0], a[1], a[2], b_hn[None, :, None, None])
    # the same value as GRUCell with a zero hidden state (functional.py:95-105): n + z * (0 - n)
    assert (h - (n + z * (0 - n))).abs().max().item() < 1e-15
    dh = _rand(tuple(h.shape), 14)
    ref = torch.autograd.grad(h, a + [b_hn], dh)
    got = T.gru_zero_hidden_gates_backward(dh, r.detach(), z.detach(), n.detach(), b_hn.detach()[None, :, None, None])
    for g_, r_ in zip(got, ref):
        assert (g_ - r_).abs().max().item() < 1e-12


def test_bn_train_forward_backward_match_torch():
    x = _rand((3, 5, 4, 6), 20).requires_grad_(True)
    gamma, beta = _rand((5,), 21).requires_grad_(True), _rand((5,), 22).requires_grad_(True)
    rm, rv = torch.zeros(5, dtype=torch.float64), torch.ones(5, dtype=torch.float64)
    ref = F.batch_norm(x, rm, rv, gamma, beta, True, 0.1, 1e-5)
    y, mean, invstd, var_unbiased = T.bn_train_forward(x.detach(), gamma.detach(), beta.detach())
    assert (y - ref).abs().max().item() < 1e-12
    # running buffers as torch updates them: momentum 0.1, unbiased variance (what restate.training_mode pins)
    assert (rm - 0.1 * mean).abs().max().item() < 1e-12 and (rv - (0.9 + 0.1 * var_unbiased)).abs().max().item() < 1e-12
    dy = _rand(tuple(ref.shape), 23)
    gx, gg, gb = torch.autograd.grad(ref, [x, gamma, beta], dy)
    dx, dgamma, dbeta = T.bn_train_backward(dy, x.detach(), mean, invstd, gamma.detach())
    assert (dx - gx).abs().max().item() < 1e-12 and (dgamma - gg).abs().max().item() < 1e-12
    assert (dbeta - gb).abs().max().item() < 1e-12


@pytest.mark.parametrize("co,ci,h,w", [(32, 5, 8, 14), (64, 3, 16, 20), (32, 2, 8, 5)])
def test_tap_packed_conv_emulation(co, ci, h, w):
    """The arithmetic of csrc/conv_pack3.cu emulated with torch: per halo position q (pitch 16, origin one pixel up-left
    of the 8 x 14 output tile) Y[q, (g, kw, co)] = sum_{kh, ci} halo[q + 16 kh, ci] * rows[(g, kw, co), ci, kh], then
    out[p] = Y[p, kw=0] + Y[p + 1, kw=1] + Y[p + 2, kw=2] -- equals conv3x3(pad 1), including partial right-edge tiles."""
    x, wt = _rand((1, ci, h, w), 30), _rand((co, ci, 3, 3), 31)
    ref = F.conv2d(x, wt, padding=1)
    rows = T.tap_packed_rows(wt)                                  # [co/32*96, ci, 3]
    g = co // 32
    out = torch.zeros_like(ref)
    for oh0 in range(0, h, 8):
        for ow0 in range(0, w, 14):
            halo = torch.zeros((ci, 10, 16), dtype=torch.float64)   # TMA box: rows oh0-1 .. oh0+8, cols ow0-1 .. ow0+14
            for r in range(10):
                for c in range(16):
                    yy, xx = oh0 - 1 + r, ow0 - 1 + c
                    if 0 <= yy < h and 0 <= xx < w:
                        halo[:, r, c] = x[0, :, yy, xx]
            flat = halo.reshape(ci, 160)
            # 128 "TMEM lanes": lane q reads halo position q + 16 kh for filter row kh
            y = torch.zeros((128, g * 96), dtype=torch.float64)
            for kh in range(3):
                a = flat[:, 16 * kh:16 * kh + 128].t()             # [128 lanes, ci]
                y += a @ rows[:, :, kh].t()                         # [128, g*96]
            y = y.view(128, g, 3, 32)
            for q in range(128):
                r, c = q >> 4, q & 15
                if c < 14 and oh0 + r < h and ow0 + c < w:
                    out[0, :, oh0 + r, ow0 + c] = (y[q, :, 0] + y[q + 1, :, 1] + y[q + 2, :, 2]).reshape(co)
    assert (out - ref).abs().max().item() < 1e-12
