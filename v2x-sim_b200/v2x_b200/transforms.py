"""Host-side weight transforms (and the element-wise backward formulas of the GRU gates / train-mode BatchNorm) that turn
three operations of the path into plain stride-1 correlations, i.e. into
launches the existing sm_100a conv kernels can run (pure index arithmetic on the reference-format OIHW weights; each
identity is proven against torch on CPU in tests/test_transforms.py).  They are the operand-preparation half of

  * the parity-decomposed decoder convs (DESIGN 3.1b "next lever"): ``conv3x3(nearest_up2(x))`` -- what
    Backbone.py:173-237 computes for the first source of conv5_1 .. conv8_1 -- equals, on each of the four output parity
    classes, a 2x2 correlation of the UN-upsampled ``x`` with pre-summed filter taps: 4/9 of the FLOPs and a quarter of
    the reads for that source;
  * the backward pass: the data gradient of a stride-1 3x3 conv is a stride-1 3x3 correlation of ``dy``
    with the transposed, 180-degree-rotated filter, and the data gradient of a stride-2 3x3 conv splits into four
    stride-1 correlations of ``dy`` (one per input parity class) -- no zero-insertion, no wasted MACs.

Nothing here launches a kernel; the GPU plans do not use these yet.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

# For output row 2i + p of conv3x3(pad 1) over a nearest-2x-upsampled map, filter row kh reads upsampled row
# 2i + p + kh - 1, i.e. source row i + floor((p + kh - 1) / 2):  p = 0: kh 0 -> i-1, kh 1,2 -> i;  p = 1: kh 0,1 -> i,
# kh 2 -> i+1.  So each parity uses two source rows, offsets (-1, 0) for p = 0 and (0, +1) for p = 1.
_UP_TAPS = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}   # parity -> (filter rows summed into the first / second source row)
_UP_OFFSETS = {0: (-1, 0), 1: (0, 1)}               # parity -> source-row offsets of those two taps


def upsample_conv_parity_weights(w: torch.Tensor) -> Dict[Tuple[int, int], torch.Tensor]:
    """OIHW 3x3 weights ``w`` -> {(py, px): [O, I, 2, 2]} such that for ``y = conv2d(nearest_up2(x), w, padding=1)``

        y[:, :, 2i + py, 2j + px] = sum_{a, b in {0,1}} w_eff[(py, px)][:, :, a, b] * x[:, :, i + dy_a, j + dx_b]

    with (dy_0, dy_1) = ``UP_OFFSETS[py]``, (dx_0, dx_1) = ``UP_OFFSETS[px]`` and zero outside the map."""
    assert w.dim() == 4 and tuple(w.shape[2:]) == (3, 3)
    out = {}
    for py in (0, 1):
        rows = [w[:, :, list(taps), :].sum(2) for taps in _UP_TAPS[py]]           # 2 x [O, I, 3]
        for px in (0, 1):
            eff = torch.stack([torch.stack([r[:, :, list(taps)].sum(2) for taps in _UP_TAPS[px]], -1) for r in rows], -2)
            out[(py, px)] = eff.contiguous()                                      # [O, I, 2(row tap), 2(col tap)]
    return out


def upsample_conv_parity_apply(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Reference evaluation of the decomposition with torch ops (used by the tests and as executable documentation):
    four 2x2 correlations of ``x`` interleaved into the [2H, 2W] output."""
    n, _, h, wd = x.shape
    weff = upsample_conv_parity_weights(w)
    y = x.new_zeros((n, w.shape[0], 2 * h, 2 * wd))
    for (py, px), we in weff.items():
        # pad so that tap a reads source row i + offsets[a]: offsets (-1, 0) need one row of zeros on top, (0, +1) below
        pt, pb = (1, 0) if py == 0 else (0, 1)
        pl, pr = (1, 0) if px == 0 else (0, 1)
        xp = torch.nn.functional.pad(x, (pl, pr, pt, pb))
        y[:, :, py::2, px::2] = torch.nn.functional.conv2d(xp, we)
    return y


def tap_packed_rows(w: torch.Tensor) -> torch.Tensor:
    """OIHW 3x3 weights with 32 | O -> the row arrangement of the tap-packed conv kernel (csrc/conv_pack3.cu):
    [O/32 * 96, I, 3] with row = (group, kw, co) and the last dim = kh.  With Y[q, (g, kw, co)] = sum_{kh, ci}
    x[q + kh * PITCH, ci] * rows[(g, kw, co), ci, kh] the conv output is out[p, 32 g + co] = sum_kw Y[p + kw, (g, kw, co)]
    (p, q linear positions in a halo whose origin is one pixel up-left of the output tile)."""
    o, i = w.shape[0], w.shape[1]
    assert w.dim() == 4 and tuple(w.shape[2:]) == (3, 3) and o % 32 == 0
    return w.view(o // 32, 32, i, 3, 3).permute(0, 4, 1, 2, 3).reshape(3 * o, i, 3).contiguous()


def dgrad_weights_stride1(w: torch.Tensor) -> torch.Tensor:
    """OIHW weights of ``y = conv2d(x, w, padding=k//2)`` (stride 1, odd k) -> weights ``wt`` [I, O, k, k] with
    ``dL/dx = conv2d(dL/dy, wt, padding=k//2)``: input / output channels swapped, filter rotated by 180 degrees."""
    assert w.dim() == 4 and w.shape[2] == w.shape[3] and w.shape[2] % 2 == 1
    return w.transpose(0, 1).flip(2, 3).contiguous()


# y[o, i, j] = sum_{kh, kw} w[kh, kw] * x[2i + kh - 1, 2j + kw - 1] (3x3, stride 2, pad 1).  Input row r = 2m + q
# receives dy row i through filter row kh = r - 2i + 1:  q = 0 (even rows): kh = 1 with i = m;  q = 1 (odd rows):
# kh = 2 with i = m, kh = 0 with i = m + 1.
_S2_TAPS = {0: ((1, 0),), 1: ((2, 0), (0, 1))}      # input parity -> ((filter index, dy offset), ...)


def dgrad_parity_weights_stride2(w: torch.Tensor) -> Dict[Tuple[int, int], Tuple[torch.Tensor, Tuple[int, ...], Tuple[int, ...]]]:
    """OIHW 3x3 weights of a stride-2, pad-1 conv -> {(qy, qx): (wt [I, O, ny, nx], row offsets, col offsets)} with

        dL/dx[:, :, 2m + qy, 2n + qx] = sum_{a, b} wt[:, :, a, b] * dL/dy[:, :, m + row_off[a], n + col_off[b]]

    (zero outside dy): 1, 2, 2 and 4 taps for the four input parity classes -- 9 MACs per dy element in total, the same
    as the forward conv, instead of the 36 a zero-inserted 3x3 correlation would spend."""
    assert w.dim() == 4 and tuple(w.shape[2:]) == (3, 3)
    wt = w.transpose(0, 1)                                                        # [I, O, kh, kw]
    out = {}
    for qy in (0, 1):
        for qx in (0, 1):
            kh = [t[0] for t in _S2_TAPS[qy]]
            kw = [t[0] for t in _S2_TAPS[qx]]
            sel = wt[:, :, kh, :][:, :, :, kw].contiguous()
            out[(qy, qx)] = (sel, tuple(t[1] for t in _S2_TAPS[qy]), tuple(t[1] for t in _S2_TAPS[qx]))
    return out


def dgrad_stride2_apply(dy: torch.Tensor, w: torch.Tensor, in_hw: Tuple[int, int]) -> torch.Tensor:
    """Reference evaluation of the stride-2 data gradient through the four parity sub-correlations (tests / docs)."""
    n, _, ho, wo = dy.shape
    h, wd = in_hw
    assert h == 2 * ho and wd == 2 * wo, "the path's stride-2 convs halve even-sized maps (Backbone.py:106-135)"
    dx = dy.new_zeros((n, w.shape[1], h, wd))
    for (qy, qx), (wt, ro, co) in dgrad_parity_weights_stride2(w).items():
        acc = 0
        for a, r in enumerate(ro):
            for b, c in enumerate(co):
                shifted = torch.nn.functional.pad(dy, (0, c, 0, r))[:, :, r:r + ho, c:c + wo]   # dy[m + r, n + c], zero beyond
                acc = acc + torch.einsum("io,nohw->nihw", wt[:, :, a, b], shifted)
        dx[:, :, qy::2, qx::2] = acc
    return dx


# ---------------------------------------------------------------------------------------------------------------------
# Element-wise backward formulas the future backward epilogues will evaluate (proven against autograd on CPU)
# ---------------------------------------------------------------------------------------------------------------------
def gru_zero_hidden_gates(a_r, a_z, a_n, b_hn):
    """Forward of the zero-hidden ConvGRU gates as the EPI_GRU epilogue computes them (functional.py:95-105 with h = 0):
    a_* are the gate pre-activations (W_ih conv + b_ih [+ b_hh for r, z]), b_hn the hidden bias of the n gate,
    broadcast over pixels.  Returns (h', r, z, n)."""
    r, z = torch.sigmoid(a_r), torch.sigmoid(a_z)
    n = torch.tanh(a_n + r * b_hn)
    return (1.0 - z) * n, r, z, n


def gru_zero_hidden_gates_backward(dh, r, z, n, b_hn):
    """d(loss)/d(a_r, a_z, a_n) and d(loss)/d(b_hn) (summed over everything but the channel dim, which is dim 1) from
    d(loss)/d(h') and the saved gate values: what the GRU conv's backward epilogue hands to its dgrad / wgrad GEMMs."""
    dn = dh * (1.0 - z)
    da_z = -dh * n * z * (1.0 - z)
    da_n = dn * (1.0 - n * n)
    da_r = da_n * b_hn * r * (1.0 - r)
    reduce_dims = [d for d in range(dh.dim()) if d != 1]
    db_hn = (da_n * r).sum(reduce_dims)
    return da_r, da_z, da_n, db_hn


def bn_train_forward(x, gamma, beta, eps=1e-5):
    """Training-mode BatchNorm2d over (N, H, W) per channel: returns (y, mean, invstd, unbiased batch variance) -- the
    statistics a conv epilogue would accumulate as per-channel sum / sum of squares."""
    dims = [0, 2, 3]
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    invstd = torch.rsqrt(var + eps)
    y = (x - mean[None, :, None, None]) * (invstd * gamma)[None, :, None, None] + beta[None, :, None, None]
    m = x.numel() // x.shape[1]
    return y, mean, invstd, var * (m / max(m - 1, 1))


def bn_train_backward(dy, x, mean, invstd, gamma):
    """(dx, dgamma, dbeta) of training-mode BatchNorm2d from the saved batch statistics:
    dx = gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat))."""
    dims = [0, 2, 3]
    xhat = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
    dbeta = dy.sum(dims)
    dgamma = (dy * xhat).sum(dims)
    m = x.numel() // x.shape[1]
    dx = (gamma * invstd)[None, :, None, None] * (dy - dbeta[None, :, None, None] / m - xhat * dgamma[None, :, None, None] / m)
    return dx, dgamma, dbeta
