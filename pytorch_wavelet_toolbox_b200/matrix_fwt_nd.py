"""Separable 2-D / 3-D boundary-wavelet transforms: ``MatrixWavedec2/3`` and ``MatrixWaverec2/3``.

SURVEY.md section 8(f) row 2.  The reference applies the 1-D orthogonalised level operator of
``MatrixWavedec`` along every axis (``separable=True``, its default):

* 2-D analysis   ``/root/reference/src/ptwt/matmul_transform_2.py:368-531`` (pad odd extents, operator along the
  width, operator along the height, split into quadrants ``ll | lh / hl | hh``)
* 2-D synthesis  ``matmul_transform_2.py:740-856``
* 3-D analysis   ``matmul_transform_3.py:131-300``; synthesis ``matmul_transform_3.py:303-480``

Here every per-axis product is one launch of ``wt_matrix_axis_fwd`` / ``wt_matrix_axis_inv`` (include/wtb200.h):
band filter + dense corner blocks along an arbitrary axis of a contiguous tensor, coalesced along the
innermost index -- no transposed copies, no sparse matrices.  The odd-extent padding sample is produced
inside the kernel.

Not built: the NON-separable 2-D operator (``separable=False``: Kronecker product of the two 1-D
operators followed by a dense QR of its boundary rows, ``matmul_transform_2.py:57-230``,
``sparse_math.py:408-587``) -- ``NotImplementedError`` says so.
"""
from __future__ import annotations

import sys
from typing import Any, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from ._shape import AxisHint, Fold, check_dtype, check_mode, check_tensor, ensure_axes, fold, unfold
from ._wavelets import as_wavelet, taps_in_dtype, filter_bank
from .constants import WaveletDetailTuple2d
from .fwt import _compute_device, _dtype_code, _no_autograd, _same_device_dtype
from .matrix_fwt import _ORTH_METHODS, _analysis_taps, _deprecated_alias, _level_blocks, _synthesis_taps

KEYS_3D = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")


class _AxisBlocks:
    """Device-resident boundary blocks of the 1-D level operator, one entry per operator size."""

    def __init__(self):
        self._cache: dict = {}

    def get(self, lo: np.ndarray, hi: np.ndarray, dtype, n: int, method: str, dev: torch.device):
        key = (lo.tobytes(), hi.tobytes(), dtype, int(n), method, str(dev))
        hit = self._cache.get(key)
        if hit is None:
            b = _level_blocks(lo, hi, dtype, int(n), method)
            flat = b.flat()
            if flat.numel() == 0:
                flat = torch.zeros(1, dtype=dtype)
            hit = (b, flat.to(dev))
            self._cache[key] = hit
        return hit


def _level_sizes(shape: Sequence[int], filt_len: int, level: int, ndim: int):
    """Operator sizes per level (even extents) and which axes were padded; stops with the reference's
    warning when an extent drops below the filter length (matmul_transform_2.py:381-393)."""
    cur = list(shape)
    sizes, pads = [], []
    for lv in range(1, level + 1):
        if any(c < filt_len for c in cur):
            what = "height and width" if ndim == 2 else "depth, height, and width"
            cur_s = ", ".join(str(c) for c in cur) if ndim == 2 else f"{cur[0]}, {cur[1]},{cur[2]}"
            sys.stderr.write(
                f"Warning: The selected number of decomposition levels {level}"
                f" is too large for the given input shape {tuple(shape)}"
                f". At level {lv}, at least one of the current signal "
                f"{what} ({cur_s}) is smaller "
                f"then the filter length {filt_len}. Therefore, the transformation "
                f"is only computed up to the decomposition level {lv - 1}.\n"
            )
            break
        pad = [c % 2 != 0 for c in cur]
        cur = [c + 1 if p else c for c, p in zip(cur, pad)]
        sizes.append(tuple(cur))
        pads.append(tuple(pad))
        cur = [c // 2 for c in cur]
    return sizes, pads


def _outer_stride(t: torch.Tensor, axis: int) -> int:
    """Stride between consecutive ``outer`` indices of a contiguous tensor viewed as [outer, n, inner]."""
    return t.shape[axis] * t.stride(axis)


class _SeparableMatrixDec:
    """Common part of MatrixWavedec2 / MatrixWavedec3."""

    _ndim = 2

    def _init(self, wavelet, level, axes, orthogonalization, odd_coeff_padding_mode):
        self.wavelet = as_wavelet(wavelet)
        self.axes = ensure_axes(axes, self._ndim)
        self.level = level
        self.orthogonalization = orthogonalization
        self.odd_coeff_padding_mode = odd_coeff_padding_mode
        self.input_signal_shape: Optional[tuple[int, ...]] = None
        self.pad_list: list[tuple[bool, ...]] = []
        self.size_list: list[tuple[int, ...]] = []
        self.padded = False
        self._blocks = _AxisBlocks()
        if self.orthogonalization not in _ORTH_METHODS:
            raise NotImplementedError
        dec_lo, dec_hi, rec_lo, rec_hi = filter_bank(self.wavelet)
        if len(dec_lo) != len(rec_lo):
            raise ValueError("All filters must have the same length")

    def _transform(self, input_signal: torch.Tensor):
        nd = self._ndim
        check_tensor(input_signal)
        check_dtype(input_signal)
        x, f = fold(input_signal, nd, self.axes)
        shape = tuple(x.shape[1:])
        if self.input_signal_shape != shape:
            self.input_signal_shape = shape
        if self.level is None:
            wlen = len(self.wavelet)
            self.level = int(np.min([np.log2(s / (wlen - 1)) for s in shape]))
        elif self.level <= 0:
            raise ValueError("level must be a positive integer.")
        lo_t, hi_t = _analysis_taps(self.wavelet, x.dtype)
        filt_len = len(lo_t)
        self.size_list, self.pad_list = _level_sizes(shape, filt_len, self.level, nd)
        self.padded = any(any(p) for p in self.pad_list)
        if self.padded:
            check_mode(self.odd_coeff_padding_mode)
        _no_autograd(input_signal, wavelet=self.wavelet)
        dev = _compute_device(x)
        on_host = not x.is_cuda
        dt = x.dtype
        odd_mode = N.MODES[self.odd_coeff_padding_mode if self.padded else "zero"]
        levels_out = []
        with torch.cuda.device(dev):
            cur = x.to(dev, non_blocking=True) if on_host else x
            stream = torch.cuda.current_stream(dev).cuda_stream
            for sizes, pads in zip(self.size_list, self.pad_list):
                cur = cur.contiguous()
                # innermost axis first, like the reference (width, then height, then depth)
                for a in range(nd, 0, -1):
                    n = sizes[a - 1]
                    blocks, flat = self._blocks.get(lo_t, hi_t, dt, n, self.orthogonalization, dev)
                    oshape = list(cur.shape)
                    oshape[a] = n
                    dst = torch.empty(oshape, dtype=dt, device=dev)
                    shape_now = cur.shape
                    outer = int(np.prod(shape_now[:a], dtype=np.int64))
                    inner = int(np.prod(shape_now[a + 1:], dtype=np.int64))
                    lib = N.load()
                    lo_arr, lo_p = N.f64_array(lo_t)
                    hi_arr, hi_p = N.f64_array(hi_t)
                    rc = lib.wt_matrix_axis_fwd(
                        _dtype_code(dt), filt_len, lo_p, hi_p, n, 1 if pads[a - 1] else 0, odd_mode, blocks.nb_top,
                        blocks.nb_bot, blocks.w_left, blocks.w_right, flat.data_ptr(), cur.data_ptr(), outer, inner,
                        _outer_stride(cur, a), cur.stride(a), dst.data_ptr(), _outer_stride(dst, a), dst.stride(a), stream)
                    N.check(rc, "wt_matrix_axis_fwd")
                    cur = dst
                levels_out.append(cur)
                half = tuple(s // 2 for s in sizes)
                cur = cur[(slice(None),) + tuple(slice(0, h) for h in half)]
            approx = cur
            if on_host:
                approx = approx.cpu()
                levels_out = [t.cpu() for t in levels_out]
        return approx, levels_out, f


class MatrixWavedec2(_SeparableMatrixDec):
    """Separable 2-D boundary-wavelet FWT (reference matmul_transform_2.py:244-567).

    Returns ``(ll, (lh, hl, hh)_n, ..., (lh, hl, hh)_1)`` exactly as the reference does: ``lh`` is low
    along the height and high along the width.
    """

    _ndim = 2

    @_deprecated_alias(boundary="orthogonalization")
    def __init__(self, wavelet: Any, level: Optional[int] = None, *, axes: AxisHint = None,
                 orthogonalization: str = "qr", separable: bool = True, odd_coeff_padding_mode: str = "zero") -> None:
        self.separable = separable
        self._init(wavelet, level, axes, orthogonalization, odd_coeff_padding_mode)

    @property
    def sparse_fwt_operator(self) -> torch.Tensor:
        # the reference offers the operator only for separable=False (matmul_transform_2.py:343-347)
        raise NotImplementedError

    def __call__(self, input_signal: torch.Tensor):
        if not self.separable:
            raise NotImplementedError(
                "the non-separable 2-D boundary operator (Kronecker product + dense QR, reference "
                "matmul_transform_2.py:57-230) is outside this package's scope; use separable=True")
        approx, levels_out, f = self._transform(input_signal)
        result: list[Any] = [unfold(approx, f)]
        for full, sizes in zip(reversed(levels_out), reversed(self.size_list)):
            h2, w2 = sizes[0] // 2, sizes[1] // 2
            lh = full[:, :h2, w2:]
            hl = full[:, h2:, :w2]
            hh = full[:, h2:, w2:]
            result.append(WaveletDetailTuple2d(unfold(lh, f), unfold(hl, f), unfold(hh, f)))
        return tuple(result)


class MatrixWavedec3(_SeparableMatrixDec):
    """Separable 3-D boundary-wavelet FWT (reference matmul_transform_3.py:66-300): ``(lll, {aad..ddd}_n, ...)``."""

    _ndim = 3

    @_deprecated_alias(boundary="orthogonalization")
    def __init__(self, wavelet: Any, level: Optional[int] = None, *, axes: AxisHint = None,
                 orthogonalization: str = "qr", odd_coeff_padding_mode: str = "zero") -> None:
        self._init(wavelet, level, axes, orthogonalization, odd_coeff_padding_mode)

    def __call__(self, input_signal: torch.Tensor):
        approx, levels_out, f = self._transform(input_signal)
        result: list[Any] = [unfold(approx, f)]
        for full, sizes in zip(reversed(levels_out), reversed(self.size_list)):
            half = [s // 2 for s in sizes]
            d = {}
            for key in KEYS_3D:
                sl = tuple(slice(half[a], None) if key[a] == "d" else slice(0, half[a]) for a in range(3))
                d[key] = unfold(full[(slice(None),) + sl], f)
            result.append(d)
        return tuple(result)


class _SeparableMatrixRec:
    _ndim = 2

    def _init(self, wavelet, axes, orthogonalization):
        self.wavelet = as_wavelet(wavelet)
        self.axes = ensure_axes(axes, self._ndim)
        self.orthogonalization = orthogonalization
        self.input_signal_shape: Optional[tuple[int, ...]] = None
        self.level: Optional[int] = None
        self.padded = False
        self._blocks = _AxisBlocks()
        if self.orthogonalization not in _ORTH_METHODS:
            raise NotImplementedError
        dec_lo, dec_hi, rec_lo, rec_hi = filter_bank(self.wavelet)
        if len(dec_lo) != len(rec_lo):
            raise ValueError("All filters must have the same length")

    def _reconstruct(self, approx: torch.Tensor, levels_in: list[dict], f: Fold) -> torch.Tensor:
        """approx [B, ..]; levels_in coarsest first, each {sub-block key -> tensor [B, ..]} without the 'a..a' block."""
        nd = self._ndim
        level = len(levels_in)
        self.level = level
        first_key = "d" * nd
        shape = tuple(2 * c for c in levels_in[-1][first_key].shape[1:]) if level else tuple(approx.shape[1:])
        self.input_signal_shape = shape
        if level == 0:
            return approx
        dt = approx.dtype
        lo_t, hi_t = _synthesis_taps(self.wavelet, dt)
        _, _, rec_lo, rec_hi = filter_bank(self.wavelet)
        rlo, rhi = taps_in_dtype(rec_lo, dt), taps_in_dtype(rec_hi, dt)
        filt_len = len(rlo)
        # the operator sizes the reference would build (matmul_transform_2.py:684-727): walk down from the
        # full extent; a level whose extent is below the filter length does not exist
        sizes, pads = _level_sizes(shape, filt_len, level, nd)
        self.padded = any(any(p) for p in pads)
        dev = _compute_device(approx)
        on_host = not approx.is_cuda
        tensors = [approx] + [t for lv in levels_in for t in lv.values()]
        _no_autograd(*tensors, wavelet=self.wavelet)
        lib = N.load()
        lo_arr, lo_p = N.f64_array(rlo)
        hi_arr, hi_p = N.f64_array(rhi)
        with torch.cuda.device(dev):
            cur = approx.to(dev, non_blocking=True) if on_host else approx
            stream = torch.cuda.current_stream(dev).cuda_stream
            for c_pos, bands in enumerate(levels_in):
                li = level - 1 - c_pos
                if li >= len(sizes):
                    raise IndexError("list index out of range")  # the reference indexes past its operator list
                dshape = tuple(bands[first_key].shape[1:])
                n_full = tuple(2 * c for c in dshape)
                if n_full != sizes[li]:
                    raise RuntimeError(
                        f"size mismatch: level operator is built for {sizes[li]} but the coefficients give {n_full}")
                for a in range(nd):
                    if cur.shape[1 + a] not in (dshape[a], dshape[a] + 1):
                        raise ValueError("All coefficients on each level must have the same shape")
                full = torch.empty((cur.shape[0],) + n_full, dtype=dt, device=dev)
                # undo the analysis padding: keep the part of the running reconstruction the details cover
                full[(slice(None),) + tuple(slice(0, c) for c in dshape)].copy_(
                    cur[(slice(None),) + tuple(slice(0, c) for c in dshape)])
                for key, t in bands.items():
                    sl = tuple(slice(dshape[a], None) if key[a] == "d" else slice(0, dshape[a]) for a in range(nd))
                    full[(slice(None),) + sl].copy_(t.to(dev, non_blocking=True) if on_host else t)
                cur = full
                for a in range(nd, 0, -1):
                    n = n_full[a - 1]
                    blocks, flat = self._blocks.get(lo_t, hi_t, dt, n, self.orthogonalization, dev)
                    dst = torch.empty_like(cur)
                    outer = int(np.prod(cur.shape[:a], dtype=np.int64))
                    inner = int(np.prod(cur.shape[a + 1:], dtype=np.int64))
                    rc = lib.wt_matrix_axis_inv(
                        _dtype_code(dt), filt_len, lo_p, hi_p, n, n, blocks.nb_top, blocks.nb_bot, blocks.w_left,
                        blocks.w_right, flat.data_ptr(), cur.data_ptr(), outer, inner, _outer_stride(cur, a), cur.stride(a),
                        dst.data_ptr(), _outer_stride(dst, a), dst.stride(a), stream)
                    N.check(rc, "wt_matrix_axis_inv")
                    cur = dst
            if on_host:
                cur = cur.cpu()
        return cur


class MatrixWaverec2(_SeparableMatrixRec):
    """Separable 2-D inverse boundary-wavelet FWT (reference matmul_transform_2.py:570-856)."""

    _ndim = 2

    @_deprecated_alias(boundary="orthogonalization")
    def __init__(self, wavelet: Any, *, axes: AxisHint = None, orthogonalization: str = "qr",
                 separable: bool = True) -> None:
        self.separable = separable
        self._init(wavelet, axes, orthogonalization)

    @property
    def sparse_ifwt_operator(self) -> torch.Tensor:
        raise NotImplementedError

    def __call__(self, coefficients) -> torch.Tensor:
        if not self.separable:
            raise NotImplementedError(
                "the non-separable 2-D boundary operator is outside this package's scope; use separable=True")
        lead = check_tensor(coefficients[0])
        check_dtype(lead)
        for el in coefficients[1:]:
            if not isinstance(el, tuple) or len(el) != 3:
                raise ValueError(
                    f"Unexpected detail coefficient type: {type(el)}. Detail coefficients must be a 3-tuple of "
                    "tensors as returned by MatrixWavedec2.")
        approx, f = fold(lead, 2, self.axes)
        levels_in = []
        prev_shape = tuple(approx.shape[1:])
        flat = [approx]
        for el in coefficients[1:]:
            lh, hl, hh = (fold(t, 2, self.axes, lead=f)[0] for t in el)
            flat += [lh, hl, hh]
            levels_in.append({"ad": lh, "da": hl, "dd": hh})
        _same_device_dtype(flat)
        cur_shape = prev_shape
        for i, bands in enumerate(levels_in):
            for t in bands.values():
                if tuple(t.shape[1:]) != cur_shape or t.shape[0] != approx.shape[0]:
                    raise ValueError("All coefficients on each level must have the same shape")
            if i + 1 < len(levels_in):
                cur_shape = tuple(levels_in[i + 1]["ad"].shape[1:])
        return unfold(self._reconstruct(approx, levels_in, f), f)


class MatrixWaverec3(_SeparableMatrixRec):
    """Separable 3-D inverse boundary-wavelet FWT (reference matmul_transform_3.py:303-480)."""

    _ndim = 3

    @_deprecated_alias(boundary="orthogonalization")
    def __init__(self, wavelet: Any, *, axes: AxisHint = None, orthogonalization: str = "qr") -> None:
        self._init(wavelet, axes, orthogonalization)

    def __call__(self, coefficients) -> torch.Tensor:
        lead = check_tensor(coefficients[0])
        check_dtype(lead)
        if len(coefficients) > 1 and type(coefficients[-1]) is not dict:
            raise ValueError("Waverec3 expects dicts of tensors.")
        approx, f = fold(lead, 3, self.axes)
        levels_in = []
        flat = [approx]
        for el in coefficients[1:]:
            if not isinstance(el, dict) or len(el) != 7:
                raise ValueError(
                    f"Unexpected detail coefficient type: {type(el)}. Detail coefficients must be a dict containing "
                    "7 tensors as returned by MatrixWavedec3.")
            bands = {k: fold(el[k], 3, self.axes, lead=f)[0] for k in KEYS_3D}
            shapes = {tuple(t.shape) for t in bands.values()}
            if len(shapes) != 1:
                raise ValueError("All coefficients on each level must have the same shape")
            flat += list(bands.values())
            levels_in.append(bands)
        _same_device_dtype(flat)
        return unfold(self._reconstruct(approx, levels_in, f), f)
