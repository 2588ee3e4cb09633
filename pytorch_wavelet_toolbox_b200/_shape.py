"""Host-side shape contract of the hot path (no arithmetic, no device work).

Re-states, in this package's own terms, the pre/post-processing semantics of the reference
(``/root/reference/src/ptwt/_util.py``): dtype gate (:545-547), axis normalisation
(:817-827, :302-304), moving the transformed axes last (:351-370), adding a missing batch
dimension or folding several leading ones (:556-564, :271-291), and undoing all of that on
the results (:613-676).  Errors are raised here, on the host, before any kernel launch.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import torch

from .constants import SUPPORTED_DTYPES

AxisHint = Union[int, Sequence[int], None]

#: ptwt mode name -> the torch padding the reference uses for it (reference _util.py:36-44)
MODE_TO_TORCH = {
    "constant": "replicate",
    "zero": "constant",
    "reflect": "reflect",
    "periodic": "circular",
    "symmetric": "symmetric",
}


def check_mode(mode: Optional[str]) -> str:
    """Validate a ptwt boundary-mode string (reference _util.py:50-68); None means reflect."""
    if mode is None:
        return "reflect"
    if mode in MODE_TO_TORCH:
        return mode
    raise ValueError(f"Padding mode not supported: {mode}")


def default_axes(n: int) -> tuple[int, ...]:
    if n < 1:
        raise ValueError(f"only natural number dimensions are allowed. given: {n}")
    return tuple(range(-n, 0))


def ensure_axes(axes: AxisHint, dim: int) -> tuple[int, ...]:
    """Normalise the ``axis`` / ``axes`` argument (reference _util.py:817-827)."""
    if axes is None:
        return default_axes(dim)
    if isinstance(axes, int):
        if dim != 1:
            raise ValueError(f"tried passing single axis to {dim}D transform")
        return (axes,)
    if len(axes) != dim:
        raise ValueError(f"tried passing {len(axes)}D axes {axes} to {dim}D transform")
    if len(set(axes)) != len(axes):
        raise ValueError("Cant transform the same axis twice.")
    return tuple(int(a) for a in axes)


def check_dtype(t: torch.Tensor) -> None:
    if t.dtype not in SUPPORTED_DTYPES:
        raise ValueError(f"Input dtype {t.dtype} not supported")


def check_tensor(obj) -> torch.Tensor:
    if not isinstance(obj, torch.Tensor):
        raise ValueError("First element of coeffs must be the approximation coefficient tensor.")
    return obj


@dataclass
class Fold:
    """How one tensor was brought to ``[batch, d1..dN]`` and how to take results back."""

    ndim: int
    axes: tuple[int, ...]
    shape_after_swap: list[int]  # "ds" in the reference

    @property
    def swapped(self) -> bool:
        return self.axes != default_axes(self.ndim)


def _perm_to_back(axes: Sequence[int], rank: int) -> list[int]:
    norm = [a + rank if a < 0 else a for a in axes]
    if len(set(norm)) != len(norm):
        raise ValueError("Cant transform the same axis twice.")
    for a in norm:
        if a < 0 or a >= rank:
            raise ValueError(f"axis {a} is out of range for a tensor with {rank} dimensions")
    front = [a for a in range(rank) if a not in norm]
    return front + norm


def move_axes_last(t: torch.Tensor, axes: Sequence[int]) -> torch.Tensor:
    return t.permute(_perm_to_back(axes, t.dim()))


def move_axes_back(t: torch.Tensor, axes: Sequence[int]) -> torch.Tensor:
    perm = _perm_to_back(axes, t.dim())
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return t.permute(inv)


def fold(t: torch.Tensor, ndim: int, axes: AxisHint, lead: Optional[Fold] = None) -> tuple[torch.Tensor, Fold]:
    """Bring ``t`` to ``[batch, d1..dN]`` (a view when possible, like the reference's reshape).

    ``lead`` (the Fold of coeffs[0]) decides batch handling for every further tensor of a
    coefficient pytree, exactly as the reference keys it on ``coeffs[0].shape`` (_util.py:556).
    """
    if ndim <= 0:
        raise ValueError("Number of dimensions must be positive")
    ax = ensure_axes(axes, ndim)
    if ax != default_axes(ndim):
        t = move_axes_last(t, ax)
    if lead is None:
        ds = list(t.shape)
        if len(ds) < ndim:
            raise ValueError(f"At least {ndim} input dimensions required.")
        lead = Fold(ndim, ax, ds)
    rank = len(lead.shape_after_swap)
    if rank == ndim:
        t = t.unsqueeze(0)
    elif rank > ndim + 1:
        t = t.reshape([math.prod(t.shape[:-ndim])] + list(t.shape[-ndim:]))
    return t, lead


def unfold(t: torch.Tensor, f: Fold) -> torch.Tensor:
    """Inverse of :func:`fold` for a result tensor ``[batch, c1..cN]`` (reference _util.py:661-674)."""
    rank = len(f.shape_after_swap)
    if rank == f.ndim:
        t = t.squeeze(0)
    elif rank > f.ndim + 1:
        t = t.reshape(list(f.shape_after_swap[: -f.ndim]) + list(t.shape[-f.ndim:]))
    if f.swapped:
        t = move_axes_back(t, f.axes)
    return t


def check_pad_feasible(mode: str, extents: Sequence[int], filt_len: int) -> None:
    """Raise what ``torch.nn.functional.pad`` raises for the reference on tiny inputs.

    ``reflect`` needs pad < size, ``circular`` needs pad <= size (survey quirk 5); the custom
    symmetric padding and the other modes accept anything (reference _util.py:163-180).
    """
    padl = (2 * filt_len - 3) // 2
    for n in extents:
        padr = padl + (n % 2)
        if mode == "reflect" and max(padl, padr) >= n:
            raise RuntimeError(
                f"Padding size should be less than the corresponding input dimension, but got: "
                f"padding ({padl}, {padr}) at a dimension of size {n}"
            )
        if mode == "periodic" and max(padl, padr) > n:
            raise RuntimeError(
                f"Padding value causes wrapping around more than once: padding ({padl}, {padr}) "
                f"at a dimension of size {n}"
            )


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m
