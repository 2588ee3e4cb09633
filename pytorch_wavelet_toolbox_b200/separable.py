"""Separable-container front ends: fswavedec2/3 and fswaverec2/3 on the fused kernels.

"Next" row 1 of SURVEY.md section 8(f).  The reference computes the same Mallat pyramid as
``wavedec2/3`` with one level-1 ``wavedec`` call per axis plus transposes and reshapes
(``/root/reference/src/ptwt/separable_conv_transform.py:36-184``) and returns the details of each
level as a dict keyed by the per-axis filter path (``'ad'`` = low-pass on axis -2, high-pass on
axis -1).  The pyramid is mathematically identical to ``wavedec2/3`` (SURVEY: all bands equal at every
level), so these functions are container re-packagings of the fused transforms: no per-axis
launches, no transpose copies.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

from ._shape import AxisHint, ensure_axes
from ._wavelets import as_wavelet, filter_bank
from .constants import DETAIL_KEYS_3D, WaveletDetailTuple2d
from .fwt import wavedec2, wavedec3, waverec2, waverec3

__all__ = ["fswavedec2", "fswavedec3", "fswaverec2", "fswaverec3"]

#: insertion order of the reference's detail dicts (recursion order of _separable_conv_dwtn_)
_KEYS_2D = ("da", "ad", "dd")
_KEYS_3D = ("daa", "ada", "dda", "aad", "dad", "add", "ddd")


def _default_level(shape, filt_len: int) -> int:
    # reference separable_conv_transform.py:140-144 (float formula, not pywt's)
    return int(min(np.log2(n / (filt_len - 1)) for n in shape))


def fswavedec2(data: torch.Tensor, wavelet: Any, *, mode: str = "reflect", level: Optional[int] = None,
               axes: AxisHint = None):
    """``(cA_n, {'da','ad','dd'}_n, ..., {...}_1)`` (reference separable_conv_transform.py:187-231)."""
    ax = ensure_axes(axes, 2)
    if level is None:
        level = _default_level([data.shape[a] for a in ax], len(filter_bank(as_wavelet(wavelet))[0]))
    res = wavedec2(data, wavelet, mode=mode, level=level, axes=ax)
    out: list[Any] = [res[0]]
    for det in res[1:]:
        out.append({"da": det.horizontal, "ad": det.vertical, "dd": det.diagonal})
    return tuple(out)


def fswavedec3(data: torch.Tensor, wavelet: Any, *, mode: str = "reflect", level: Optional[int] = None,
               axes: AxisHint = None):
    """``(cA_n, {7 keys}_n, ..., {...}_1)`` (reference separable_conv_transform.py:234-278)."""
    ax = ensure_axes(axes, 3)
    if level is None:
        level = _default_level([data.shape[a] for a in ax], len(filter_bank(as_wavelet(wavelet))[0]))
    res = wavedec3(data, wavelet, mode=mode, level=level, axes=ax)
    out: list[Any] = [res[0]]
    for det in res[1:]:
        out.append({k: det[k] for k in _KEYS_3D})
    return tuple(out)


def _check(coeffs) -> None:
    if not isinstance(coeffs[0], torch.Tensor):
        raise ValueError("approximation tensor must be first in coefficient list.")
    if not all(isinstance(c, dict) for c in coeffs[1:]):
        raise ValueError("All entries after approximation tensor must be dicts.")


def _crop_like(t: torch.Tensor, ref: torch.Tensor, ndim: int) -> torch.Tensor:
    """The reference crops the running approximation to the detail's extents before each level
    (separable_conv_transform.py:95: "undo any analysis padding")."""
    if t.shape[-ndim:] == ref.shape[-ndim:]:
        return t
    sl = (Ellipsis,) + tuple(slice(0, s) for s in ref.shape[-ndim:])
    return t[sl]


def fswaverec2(coeffs, wavelet: Any, *, axes: AxisHint = None) -> torch.Tensor:
    """Inverse of :func:`fswavedec2` (reference separable_conv_transform.py:281-313)."""
    _check(coeffs)
    ax = ensure_axes(axes, 2)
    seq: list[Any] = [coeffs[0]]
    for d in coeffs[1:]:
        seq.append(WaveletDetailTuple2d(d["da"], d["ad"], d["dd"]))
    if len(seq) > 1 and ax == (-2, -1):
        seq[0] = _crop_like(seq[0], seq[1][0], 2)
    return waverec2(tuple(seq), wavelet, axes=ax)


def fswaverec3(coeffs, wavelet: Any, *, axes: AxisHint = None) -> torch.Tensor:
    """Inverse of :func:`fswavedec3` (reference separable_conv_transform.py:316-348)."""
    _check(coeffs)
    ax = ensure_axes(axes, 3)
    seq: list[Any] = [coeffs[0]]
    for d in coeffs[1:]:
        seq.append({k: d[k] for k in DETAIL_KEYS_3D})
    if len(seq) > 1 and ax == (-3, -2, -1):
        seq[0] = _crop_like(seq[0], seq[1]["aad"], 3)
    return waverec3(tuple(seq), wavelet, axes=ax)
