"""Closed-form float64 restatement of one filter-bank level -- TEST INFRASTRUCTURE ONLY.

Independent of torch's convolution kernels: plain index arithmetic in numpy, used on small cases to
check both the reference-shaped port (oracle/ptwt_port.py) and the CUDA kernels.

  analysis   c_k[i] = sum_{m<L} dec_k[m] * ext(x)[2 i + (L-1-padl) - m],   padl = (2L-3)//2
             = F.pad(padl, padl + n%2) followed by conv1d(stride=2) with the flipped filter
             (/root/reference/src/ptwt/conv_transform.py:133-141, _util.py:198-228)
  synthesis  y[t] = sum_i lo[i] rec_lo[t + padl - 2 i] + hi[i] rec_hi[t + padl - 2 i]
             = conv_transpose1d(stride=2) cropped by padl on both sides
             (/root/reference/src/ptwt/conv_transform.py:184-199)
  ext        zero / constant (edge value) / reflect (no edge repeat) / periodic / symmetric (edge
             repeated, period 2n) (/root/reference/src/ptwt/constants.py:85-110, _util.py:163-195)
"""
from __future__ import annotations

import numpy as np


def ext_index(j: int, n: int, mode: str) -> int:
    """Source index of ext(x)[j], or -1 for an implicit zero."""
    if 0 <= j < n:
        return j
    if mode == "zero":
        return -1
    if mode == "constant":
        return 0 if j < 0 else n - 1
    if mode == "reflect":
        if n == 1:
            return 0
        p = 2 * n - 2
        j %= p
        return j if j < n else p - j
    if mode == "periodic":
        return j % n
    if mode == "symmetric":
        p = 2 * n
        j %= p
        return j if j < n else p - 1 - j
    raise ValueError(f"Padding mode not supported: {mode}")


def coeff_len(n: int, L: int) -> int:
    padl = (2 * L - 3) // 2
    return (n + 2 * padl + n % 2 - L) // 2 + 1


def dwt_axis(x: np.ndarray, dec_lo, dec_hi, mode: str, axis: int = -1):
    """One analysis level along ``axis`` -> (lo, hi)."""
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    L = len(dec_lo)
    padl = (2 * L - 3) // 2
    m = coeff_len(n, L)
    lo = np.zeros(x.shape[:-1] + (m,))
    hi = np.zeros(x.shape[:-1] + (m,))
    for i in range(m):
        for k in range(L):
            s = ext_index(2 * i + k - padl, n, mode)
            if s < 0:
                continue
            lo[..., i] += dec_lo[L - 1 - k] * x[..., s]
            hi[..., i] += dec_hi[L - 1 - k] * x[..., s]
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def idwt_axis(lo: np.ndarray, hi: np.ndarray, rec_lo, rec_hi, keep: int | None = None, axis: int = -1):
    """One synthesis level along ``axis``; ``keep`` samples are returned (default 2(m-1)+L-2 padl)."""
    lo = np.moveaxis(np.asarray(lo, dtype=np.float64), axis, -1)
    hi = np.moveaxis(np.asarray(hi, dtype=np.float64), axis, -1)
    m = lo.shape[-1]
    L = len(rec_lo)
    padl = (2 * L - 3) // 2
    full = 2 * (m - 1) + L - 2 * padl
    keep = full if keep is None else keep
    y = np.zeros(lo.shape[:-1] + (keep,))
    for t in range(keep):
        u = t + padl
        for i in range(m):
            k = u - 2 * i
            if 0 <= k < L:
                y[..., t] += rec_lo[k] * lo[..., i] + rec_hi[k] * hi[..., i]
    return np.moveaxis(y, -1, axis)


def dwt_nd_level(x: np.ndarray, dec_lo, dec_hi, mode: str, ndim: int):
    """One separable N-d level over the last ``ndim`` axes -> list of 2^ndim bands, index
    k = sum_a hi(a) << (ndim-1-a) with axis 0 the slowest (include/wtb200.h)."""
    bands = [np.asarray(x, dtype=np.float64)]
    for a in range(ndim):  # slowest first; each existing band splits into (lo, hi) along axis a
        nxt = []
        for b in bands:
            lo, hi = dwt_axis(b, dec_lo, dec_hi, mode, axis=-(ndim - a))
            nxt.append((lo, hi))
        bands = [v for p in nxt for v in p]
    # after the loop the list order is lexicographic in (axis0, axis1, ...) with lo before hi
    return bands


def idwt_nd_level(bands, rec_lo, rec_hi, ndim: int, keep=None):
    """Inverse of :func:`dwt_nd_level`; ``keep`` = per-axis sample counts."""
    cur = list(bands)
    for a in range(ndim - 1, -1, -1):
        nxt = []
        for j in range(0, len(cur), 2):
            k = None if keep is None else keep[a]
            nxt.append(idwt_axis(cur[j], cur[j + 1], rec_lo, rec_hi, k, axis=-(ndim - a)))
        cur = nxt
    return cur[0]
