"""Batch sharding of the transforms over the GPUs of one box (one process per GPU).

Every batch item (any folded leading dimension) is transformed independently -- the reference has no
cross-sample reduction and no halo between samples -- so the path shards with
NO data-path collective: each rank transforms its contiguous slice of dim 0.  The only collective is
the optional collection of the results: the coefficient pytree of a rank is packed into ONE
contiguous ``[B_local, P]`` buffer (identical layout on every rank, because extents depend only on
the sample shape, the filter length and the level count) and gathered with ONE ``all_gather``.
"""
from __future__ import annotations

from typing import Any, Optional, Sequence

import torch

from .constants import DETAIL_KEYS_3D, WaveletDetailTuple2d

__all__ = ["shard_bounds", "shard", "pack_coeffs", "unpack_coeffs", "all_gather_coeffs"]


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced slices of ``range(n)``: the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def shard(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], world)[rank]
    return x[lo:hi]


def _flatten(coeffs) -> tuple[list[torch.Tensor], list[Any]]:
    tensors, spec = [coeffs[0]], ["approx"]
    for el in coeffs[1:]:
        if isinstance(el, torch.Tensor):
            tensors.append(el)
            spec.append("t")
        elif isinstance(el, dict):
            tensors.extend(el[k] for k in DETAIL_KEYS_3D)
            spec.append("d")
        else:
            tensors.extend(el)
            spec.append("h")
    return tensors, spec


def pack_coeffs(coeffs) -> tuple[torch.Tensor, dict]:
    """Coefficient pytree with a leading batch dim -> (``[B, P]`` contiguous buffer, layout meta)."""
    tensors, spec = _flatten(coeffs)
    b = tensors[0].shape[0]
    shapes = [tuple(t.shape[1:]) for t in tensors]
    flat = torch.cat([t.reshape(b, -1) for t in tensors], 1) if b or tensors else tensors[0].reshape(b, -1)
    meta = {"shapes": shapes, "spec": spec, "list": isinstance(coeffs, list)}
    return flat.contiguous(), meta


def unpack_coeffs(flat: torch.Tensor, meta: dict):
    b = flat.shape[0]
    views, off = [], 0
    for shp in meta["shapes"]:
        n = 1
        for s in shp:
            n *= s
        views.append(flat[:, off: off + n].reshape((b,) + tuple(shp)))
        off += n
    out: list[Any] = [views[0]]
    i = 1
    for kind in meta["spec"][1:]:
        if kind == "t":
            out.append(views[i]); i += 1
        elif kind == "h":
            out.append(WaveletDetailTuple2d(*views[i: i + 3])); i += 3
        else:
            out.append(dict(zip(DETAIL_KEYS_3D, views[i: i + 7]))); i += 7
    return out if meta["list"] else tuple(out)


def all_gather_coeffs(coeffs, total_batch: int, group=None):
    """Collect the shards of all ranks with ONE all_gather of the packed buffers.

    ``total_batch`` is the global batch size; shards follow :func:`shard_bounds`.  Uneven shards are
    padded to the largest one for the collective and trimmed afterwards.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    bounds = shard_bounds(total_batch, world)
    bmax = max(hi - lo for lo, hi in bounds)
    flat, meta = pack_coeffs(coeffs)
    if flat.shape[0] < bmax:
        pad = torch.zeros((bmax - flat.shape[0], flat.shape[1]), dtype=flat.dtype, device=flat.device)
        flat = torch.cat([flat, pad], 0)
    out = torch.empty((world * bmax, flat.shape[1]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    if any(hi - lo != bmax for lo, hi in bounds):
        out = torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(bounds)], 0)
    return unpack_coeffs(out, meta)
