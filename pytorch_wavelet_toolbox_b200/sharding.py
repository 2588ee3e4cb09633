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

__all__ = ["shard_bounds", "shard", "pack_coeffs", "unpack_coeffs", "all_gather_coeffs", "transform_and_gather"]


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


def _packed_base(tensors: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """The ONE ``[B, P]`` buffer the transforms of this package return views of (``fwt._make_plan``: every band a dense
    block inside a per-item record of P elements), or None for any other layout."""
    t0 = tensors[0]
    base = t0._base
    if base is None or base.dim() != 2 or not base.is_contiguous() or base.storage_offset() != 0:
        return None
    ptr = base.untyped_storage().data_ptr()
    for t in tensors:
        if t.dim() < 1 or t.shape[0] != base.shape[0] or t.untyped_storage().data_ptr() != ptr:
            return None
        if t.numel() and t.shape[0] > 1 and t.stride(0) != base.stride(0):
            return None
    return base


def _rebuild(coeffs, fn):
    """Same pytree with every tensor replaced by fn(tensor)."""
    out: list[Any] = []
    for el in coeffs:
        if isinstance(el, torch.Tensor):
            out.append(fn(el))
        elif isinstance(el, dict):
            out.append({k: fn(v) for k, v in el.items()})
        else:
            out.append(type(el)(*[fn(v) for v in el]) if hasattr(el, "_fields") else type(el)(fn(v) for v in el))
    return out if isinstance(coeffs, list) else tuple(out)


def all_gather_coeffs(coeffs, total_batch: int, group=None):
    """Collect the shards of all ranks with ONE all_gather.

    ``total_batch`` is the global batch size; shards follow :func:`shard_bounds`.  The coefficient tensors this package
    returns are views of one packed ``[B_local, P]`` buffer per call, identical in layout on every rank: that buffer IS
    the message (no packing copy), and the result is the same pytree of views over the gathered ``[B_total, P]``
    buffer.  Foreign layouts and uneven shards are packed (``pack_coeffs``) and, if uneven, padded to the largest
    shard for the collective and trimmed afterwards.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    bounds = shard_bounds(total_batch, world)
    bmax = max(hi - lo for lo, hi in bounds)
    even = all(hi - lo == bmax for lo, hi in bounds)
    tensors, _ = _flatten(coeffs)
    base = _packed_base(tensors) if even else None
    if base is not None and base.shape[0] == bmax:
        p = base.shape[1]
        out = torch.empty((world * bmax, p), dtype=base.dtype, device=base.device)
        dist.all_gather_into_tensor(out, base, group=group)
        return _rebuild(coeffs, lambda t: out.as_strided((world * bmax,) + tuple(t.shape[1:]),
                                                        (p,) + tuple(t.stride()[1:]), t.storage_offset()))
    flat, meta = pack_coeffs(coeffs)
    if flat.shape[0] < bmax:
        pad = torch.zeros((bmax - flat.shape[0], flat.shape[1]), dtype=flat.dtype, device=flat.device)
        flat = torch.cat([flat, pad], 0)
    out = torch.empty((world * bmax, flat.shape[1]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    if not even:
        out = torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(bounds)], 0)
    return unpack_coeffs(out, meta)


def transform_and_gather(transform, x_local: torch.Tensor, chunks: int = 4, group=None) -> list:
    """Transform this rank's shard chunk by chunk and collect every chunk from all ranks, the collective of chunk k
    overlapping the transform of chunk k + 1 (over NVLink the gather costs far more than the
    transform, so it is the part to hide).  ``transform`` maps ``[b, ...]`` to a coefficient pytree (e.g.
    ``lambda t: wavedec2(t, "db8", level=5)``).  Returns one gathered pytree per chunk; chunk ``k`` holds, for every
    rank ``r`` in order, the items ``shard_bounds(B_local, chunks)[k]`` of that rank's shard (the ranks' shards must
    have equal sizes)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    on_gpu = x_local.is_cuda
    results = []
    if on_gpu:
        dev = x_local.device
        cur = torch.cuda.current_stream(dev)
        comm = _comm_stream(dev)
        comm.wait_stream(cur)
    for lo, hi in shard_bounds(x_local.shape[0], max(1, min(chunks, x_local.shape[0]))):
        if hi == lo:
            continue
        c = transform(x_local[lo:hi])
        if not on_gpu:
            results.append(all_gather_coeffs(c, (hi - lo) * world, group))
            continue
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(comm):
            comm.wait_event(done)
            results.append(all_gather_coeffs(c, (hi - lo) * world, group))
            for t in _flatten(c)[0]:
                t.record_stream(comm)
    if on_gpu:
        cur.wait_stream(comm)
    return results


_comm_streams: dict = {}


def _comm_stream(dev: torch.device):
    if dev not in _comm_streams:
        _comm_streams[dev] = torch.cuda.Stream(device=dev)
    return _comm_streams[dev]
