"""Multi-GPU host logic on CPU: world_size-2 gloo, batch sharding + ONE packed all_gather."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import flatten_coeffs
from oracle import ptwt_port as P
from pytorch_wavelet_toolbox_b200 import sharding as S


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 8, 64, 4096):
        for w in (1, 2, 3, 8):
            b = S.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip_all_containers():
    x1 = torch.randn(3, 50)
    c1 = P.wavedec(x1, "db2", level=2)
    x2 = torch.randn(3, 20, 24)
    c2 = P.wavedec2(x2, "db2", level=2)
    x3 = torch.randn(3, 10, 12, 14)
    c3 = P.wavedec3(x3, "haar", level=1)
    for c in (c1, c2, c3):
        flat, meta = S.pack_coeffs(c)
        assert flat.dim() == 2 and flat.is_contiguous()
        back = S.unpack_coeffs(flat, meta)
        assert type(back) is type(c)
        for a, b in zip(flatten_coeffs(back), flatten_coeffs(c)):
            assert torch.equal(a, b)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, total: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(total, 24, 28, generator=g, dtype=torch.float64)
        local = S.shard(x, rank, world)
        # the transform of the shard (the oracle stands in for the CUDA path on this CPU-only box)
        c_local = P.wavedec2(local, "db2", level=2)
        full = S.all_gather_coeffs(c_local, total)
        want = P.wavedec2(x, "db2", level=2)
        ok = all(torch.equal(a, b) for a, b in zip(flatten_coeffs(full), flatten_coeffs(want)))
        q.put((rank, ok, [tuple(t.shape) for t in flatten_coeffs(full)][:2]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 7])
def test_two_rank_gloo_gather_equals_unsharded_transform(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert all(ok for _, ok, _ in res), res
    assert res[0][2][0][0] == total


def _packed_like_the_cuda_path(c):
    """Coefficient pytree whose tensors are views of ONE [B, P] buffer -- the layout fwt._analysis returns."""
    flat, meta = S.pack_coeffs(c)
    views = S.unpack_coeffs(flat, meta)
    assert S._packed_base(S._flatten(views)[0]) is not None
    return views


def _worker_packed(rank: int, world: int, port: int, total: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(9)
        x = torch.randn(total, 20, 24, generator=g, dtype=torch.float64)
        local = S.shard(x, rank, world)
        want = P.wavedec2(x, "db2", level=2)
        # (1) zero-copy path: the packed buffer itself is the message, the result is views of the gathered buffer
        full = S.all_gather_coeffs(_packed_like_the_cuda_path(P.wavedec2(local, "db2", level=2)), total)
        ok1 = type(full) is type(want) and all(torch.equal(a, b) for a, b in zip(flatten_coeffs(full), flatten_coeffs(want)))
        ok1 = ok1 and S._packed_base(S._flatten(full)[0]) is not None
        # (2) chunked transform + gather: chunk k holds rank-major the k-th slice of every shard
        chunks = 2
        res = S.transform_and_gather(lambda t: _packed_like_the_cuda_path(P.wavedec2(t, "db2", level=2)), local, chunks=chunks)
        ok2 = len(res) == chunks
        per = total // world
        for k, (lo, hi) in enumerate(S.shard_bounds(per, chunks)):
            idx = [r * per + i for r in range(world) for i in range(lo, hi)]
            for a, b in zip(flatten_coeffs(res[k]), flatten_coeffs(want)):
                ok2 = ok2 and torch.equal(a, b[idx])
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_zero_copy_gather_and_chunked_overlap():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert all(a and b for _, a, b in res), res
