"""Differentiable path of the padded transforms.

The reference is differentiable by construction (it is a chain of torch operators); users train
through it (e.g. ``examples/network_compression/wavelet_linear.py:118,150``).  Under grad mode the
transforms therefore switch from the fused multi-level launch to a level-by-level chain whose only
non-torch nodes are two ``torch.autograd.Function`` s around the CUDA kernels:

* one analysis level with ZERO extension, whose adjoint is exactly one synthesis level with the
  flipped decomposition filters (the crop of the transposed convolution removes the zero padding);
* one synthesis level, whose adjoint is one zero-extension analysis level with the flipped
  reconstruction filters.

The boundary extension itself (reflect / constant / periodic / symmetric) is applied beforehand with
differentiable torch indexing, so its adjoint (the fold of the halo back into the signal) is torch's;
because ``pad_left = L - 2`` is even, the level outputs are the slice ``[pad_left/2 : pad_left/2 + M]``
of the zero-extension transform of the explicitly extended signal.  Gradients w.r.t. the filter taps
(learnable wavelets) are not provided: tensors with ``requires_grad`` as filters raise.
"""
from __future__ import annotations

from typing import Any, Sequence

import torch
import torch.nn.functional as F

_TORCH_MODE = {"constant": "replicate", "zero": "constant", "reflect": "reflect", "periodic": "circular"}


def _floats(seq) -> tuple:
    if isinstance(seq, torch.Tensor):
        return tuple(float(v) for v in seq.detach().cpu().reshape(-1))
    return tuple(float(v) for v in seq)


def _sym_pad_axis(x: torch.Tensor, axis: int, left: int, right: int) -> torch.Tensor:
    n = x.shape[axis]
    if left > n or right > n:  # longer than the signal: extend in steps (reference _util.py:163-173)
        if left > n:
            x = _sym_pad_axis(x, axis, n, 0)
            left -= n
        if right > n:
            x = _sym_pad_axis(x, axis, 0, n)
            right -= n
        return _sym_pad_axis(x, axis, left, right)
    parts = [x]
    if left:
        parts.insert(0, x.narrow(axis, 0, left).flip(axis))
    if right:
        parts.append(x.narrow(axis, n - right, right).flip(axis))
    return torch.cat(parts, axis)


def extend(x: torch.Tensor, ndim: int, filt_len: int, mode: str) -> torch.Tensor:
    """Boundary extension of ``[B, d1..dN]`` by (L-2, L-2 + n%2) per axis with torch ops."""
    base = (2 * filt_len - 3) // 2
    pads = [(base, base + x.shape[1 + a] % 2) for a in range(ndim)]
    if mode == "symmetric":
        for a, (l, r) in enumerate(pads):
            x = _sym_pad_axis(x, 1 + a, l, r)
        return x
    flat: list[int] = []
    for l, r in reversed(pads):
        flat += [l, r]
    return F.pad(x.unsqueeze(1), flat, mode=_TORCH_MODE[mode]).squeeze(1)


class ZeroLevelAnalysis(torch.autograd.Function):
    """One analysis level with zero extension: ``x [B, d..] -> 2^ndim bands``."""

    @staticmethod
    def forward(ctx, x, dec_lo: tuple, dec_hi: tuple, ndim: int):
        from . import fwt

        wav = (list(dec_lo), list(dec_hi), list(dec_lo), list(dec_hi))
        approx, details, _ = fwt._analysis(x, wav, "zero", 1, None, ndim)
        ctx.taps = (dec_lo, dec_hi)
        ctx.ndim = ndim
        ctx.in_shape = tuple(x.shape)
        return (approx,) + tuple(details[0])

    @staticmethod
    def backward(ctx, *grads):
        from . import fwt

        dec_lo, dec_hi = ctx.taps
        ndim = ctx.ndim
        ref = next(g for g in grads if g is not None)
        bands = [g if g is not None else torch.zeros_like(ref) for g in grads]
        # adjoint of (zero pad -> stride-2 correlation) = transposed convolution with the same kernel,
        # cropped by the pad: the synthesis kernel with rec := flipped dec
        wav = (None, None, list(dec_lo)[::-1], list(dec_hi)[::-1])
        f = fwt.Fold(ndim, tuple(range(-ndim, 0)), list(ctx.in_shape))
        gx = fwt._synthesis(bands[0].contiguous(), [[b.contiguous() for b in bands[1:]]], [bands[1]], wav, ndim, f)
        sl = (slice(None),) + tuple(slice(0, n) for n in ctx.in_shape[1:])
        return gx[sl], None, None, None


class LevelSynthesis(torch.autograd.Function):
    """One synthesis level: ``2^ndim bands -> y [B, 2c - L + 2 ..]``."""

    @staticmethod
    def forward(ctx, rec_lo: tuple, rec_hi: tuple, ndim: int, *bands):
        from . import fwt

        wav = (None, None, list(rec_lo), list(rec_hi))
        f = fwt.Fold(ndim, tuple(range(-ndim, 0)), list(bands[0].shape))
        y = fwt._synthesis(bands[0], [list(bands[1:])], [bands[1]], wav, ndim, f)
        ctx.taps = (rec_lo, rec_hi)
        ctx.ndim = ndim
        ctx.coeff_shape = tuple(bands[0].shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import fwt

        rec_lo, rec_hi = ctx.taps
        # adjoint of (transposed convolution -> crop) = zero-extension analysis with dec := flipped rec
        wav = (list(rec_lo)[::-1], list(rec_hi)[::-1], None, None)
        approx, details, _ = fwt._analysis(gy.contiguous(), wav, "zero", 1, None, ctx.ndim)
        out = [approx] + list(details[0])
        sl = (slice(None),) + tuple(slice(0, n) for n in ctx.coeff_shape[1:])
        return (None, None, None) + tuple(t[sl] for t in out)


def analysis_with_grad(x: torch.Tensor, dec_lo, dec_hi, mode: str, level: int, ndim: int, dev: torch.device):
    """Level-by-level differentiable analysis of folded data ``[B, d1..dN]``; returns
    (approx, [coarsest-first lists of bands k = 1..])."""
    from . import _native as N
    from ._shape import check_pad_feasible

    if isinstance(dec_lo, torch.Tensor) and (dec_lo.requires_grad or dec_hi.requires_grad):
        raise NotImplementedError("gradients with respect to the filter taps are not implemented")
    lo, hi = _floats(dec_lo), _floats(dec_hi)
    L = len(lo)
    if L % 2:
        raise NotImplementedError("the differentiable path needs an even filter length")
    home = x.device
    cur = x.to(dev)
    shift = ((2 * L - 3) // 2) // 2
    details = []
    for _ in range(level):
        dims = tuple(cur.shape[1:])
        check_pad_feasible(mode, dims, L)
        m = tuple(N.coeff_len(n, L) for n in dims)
        xp = cur if mode == "zero" else extend(cur, ndim, L, mode)
        bands = ZeroLevelAnalysis.apply(xp, lo, hi, ndim)
        if mode != "zero":
            sl = (slice(None),) + tuple(slice(shift, shift + mm) for mm in m)
            bands = tuple(b[sl] for b in bands)
        details.append([b.to(home) for b in bands[1:]])
        cur = bands[0]
    details.reverse()
    return cur.to(home), details


def synthesis_with_grad(approx: torch.Tensor, levels_in, probes, rec_lo, rec_hi, ndim: int, dev: torch.device):
    """Level-by-level differentiable synthesis; arguments as fwt._synthesis (already validated)."""
    if isinstance(rec_lo, torch.Tensor) and (rec_lo.requires_grad or rec_hi.requires_grad):
        raise NotImplementedError("gradients with respect to the filter taps are not implemented")
    lo, hi = _floats(rec_lo), _floats(rec_hi)
    home = approx.device
    cur = approx.to(dev)
    for i, bands in enumerate(levels_in):
        want = tuple(bands[0].shape[1:])
        if tuple(cur.shape[1:]) != want:
            raise ValueError("All coefficients on each level must have the same shape")
        y = LevelSynthesis.apply(lo, hi, ndim, cur, *[b.to(dev) for b in bands])
        if i + 1 < len(levels_in):
            nxt = tuple(probes[i + 1].shape[1:])
            sl = [slice(None)]
            for a in range(ndim):
                if nxt[a] == y.shape[1 + a]:
                    sl.append(slice(None))
                elif nxt[a] == y.shape[1 + a] - 1:
                    sl.append(slice(0, nxt[a]))
                else:
                    raise AssertionError("padding error, please check if dec and rec wavelets are identical.")
            y = y[tuple(sl)]
        cur = y
    return cur.to(home)
