"""GPU parity of the code paths the headline benchmark actually times ("parity gaps"):

* the two-level kernel of independent warps (csrc/fused2d_wpair.cuh) on every boundary mode, filter length,
  strip / segment layout, compared image by image with the oracle;
* the chunked two-stream branch of the 2-D analysis (batch halves on the caller's and the library's stream with
  reused scratch slots), forced on small data and run at its real size;
* the host pipeline at its real chunk size.

Tolerance: |delta| <= 1e-5 * max|c| (float32), c = the oracle's coefficients of the SAME image.
"""
from __future__ import annotations

import pytest
import torch

import pytorch_wavelet_toolbox_b200 as wt
from conftest import assert_close_rel, flatten_coeffs
from oracle import ptwt_port as P
from pytorch_wavelet_toolbox_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cmp_every_image(got, x, wav, mode, level, what):
    """Compare every batch item with the oracle run on that item alone."""
    fg = flatten_coeffs(got)
    for i in range(x.shape[0]):
        want = flatten_coeffs(P.wavedec2(x[i:i + 1], wav, mode=mode, level=level))
        scale = max(float(t.abs().max()) for t in want)
        assert len(want) == len(fg)
        for j, (a, b) in enumerate(zip(fg, want)):
            assert_close_rel(a[i:i + 1], b, scale=scale, what=f"{what} image {i} tensor {j}")


@pytest.mark.parametrize("mode", ["zero", "constant", "reflect", "symmetric", "periodic"])
@pytest.mark.parametrize("wav", ["haar", "db2", "db3", "db4", "sym4"])
def test_wpair_kernel_every_mode_and_filter(knob, mode, wav):
    """Levels (1,2) and (3,4) through fwd2d_wpair_kernel (periodic falls back to one launch per level)."""
    knob("WPAIR", 1)
    knob("WPAIR_MIN", 1)
    knob("WPAIR_DEEP", 1)
    g = torch.Generator().manual_seed(5)
    for shape, level in (((2, 96, 160), 2), ((3, 132, 260), 3), ((1, 512, 1024), 4), ((2, 260, 132), 2), ((1, 64, 640), 2)):
        x = torch.randn(shape, generator=g)
        _native.launch_count_reset()
        got = wt.wavedec2(x.to(DEV), wav, mode=mode, level=level)
        n_launch = _native.launch_count()
        _cmp_every_image(got, x, wav, mode, level, f"wpair {wav} {mode} {shape} L{level}")
        if mode != "periodic":
            assert n_launch < level, f"{wav} {mode} {shape}: {n_launch} launches for {level} levels - pair kernel not used"
        rec = wt.waverec2(got, wav)
        assert_close_rel(rec[..., :shape[1], :shape[2]], P.waverec2(P.wavedec2(x, wav, mode=mode, level=level), wav)[..., :shape[1], :shape[2]],
                         scale=float(x.abs().max()), what="waverec2 of the pair kernel's output")


@pytest.mark.parametrize("seg", [8, 13, 40])
def test_wpair_segment_restarts(knob, seg):
    """Short row segments: every segment restarts both levels from its halo; the last one is ragged."""
    knob("WPAIR", 1)
    knob("WPAIR_MIN", 1)
    knob("WPAIR_SEG", seg)
    g = torch.Generator().manual_seed(6 + seg)
    for mode in ("zero", "constant", "reflect", "symmetric"):
        for shape in ((2, 300, 200), (1, 130, 512), (2, 517, 136)):
            x = torch.randn(shape, generator=g)
            got = wt.wavedec2(x.to(DEV), "db4", mode=mode, level=2)
            _cmp_every_image(got, x, "db4", mode, 2, f"wpair seg={seg} {mode} {shape}")
            got = wt.wavedec2(x.to(DEV), "db3", mode=mode, level=3)
            _cmp_every_image(got, x, "db3", mode, 3, f"wpair seg={seg} db3 {mode} {shape}")


def test_wpair_wide_and_narrow_strips(knob):
    """Strip layouts: one strip, many strips, a last strip of a few columns, a shifted last strip."""
    knob("WPAIR", 1)
    knob("WPAIR_MIN", 1)
    g = torch.Generator().manual_seed(9)
    for w in (64, 120, 236, 240, 244, 248, 252, 484, 1000, 4096):
        x = torch.randn(2, 72, w, generator=g)
        for mode in ("reflect", "symmetric", "zero", "constant"):
            got = wt.wavedec2(x.to(DEV), "db4", mode=mode, level=2)
            _cmp_every_image(got, x, "db4", mode, 2, f"wpair width {w} {mode}")


@pytest.mark.parametrize("mode", ["zero", "constant", "reflect", "periodic", "symmetric"])
def test_chunked_two_stream_branch_small(knob, mode):
    """The branch the headline times (batch cut into chunks that alternate between the caller's stream and the
    library's auxiliary stream, scratch slots reused per stream), forced on small data: batch 7, chunks of 3."""
    knob("CHUNK", 3)
    knob("WPAIR", 1)
    knob("WPAIR_MIN", 1)
    g = torch.Generator().manual_seed(31)
    for level in (2, 3, 4):
        x = torch.randn(7, 200, 264, generator=g)
        got = wt.wavedec2(x.to(DEV), "db4", mode=mode, level=level)
        torch.cuda.synchronize()
        _cmp_every_image(got, x, "db4", mode, level, f"chunked {mode} L{level}")
    with _native.knobs(WPAIR=0):
        x = torch.randn(7, 200, 264, generator=g)
        got = wt.wavedec2(x.to(DEV), "db4", mode=mode, level=3)
        _cmp_every_image(got, x, "db4", mode, 3, f"chunked per-level {mode}")


def test_headline_configuration_every_image():
    """BASELINE configs[1] at its real size: 64 x 4096^2 float32, db4, level 4, reflect -- exactly the call
    bench.py times.  Images 0, 31, 32, 63 (both batch halves, i.e. both streams and the reused scratch slots)
    against the oracle; all 64 through the round trip."""
    g = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.randn(64, 4096, 4096, generator=g, device=DEV)
    c = wt.wavedec2(x, "db4", level=4)
    flat = flatten_coeffs(c)
    for i in (0, 31, 32, 63):
        xi = x[i:i + 1].cpu()
        want = flatten_coeffs(P.wavedec2(xi, "db4", level=4))
        scale = max(float(t.abs().max()) for t in want)
        for j, (a, b) in enumerate(zip(flat, want)):
            assert_close_rel(a[i:i + 1], b, scale=scale, what=f"headline image {i} tensor {j}")
    rec = wt.waverec2(c, "db4")
    err = (rec - x).abs().amax(dim=(1, 2))
    assert float(err.max()) < 2e-5 * float(x.abs().max()), f"round trip per image: {err.tolist()}"
    del rec
    # the opt-in two-level kernel must agree with the per-level kernels on the same data
    with _native.knobs(WPAIR=1):
        c2 = flatten_coeffs(wt.wavedec2(x, "db4", level=4))
    scale = max(float(t.abs().max()) for t in c2)
    for j, (a, b) in enumerate(zip(flat, c2)):
        assert float((a - b).abs().max()) <= 1e-5 * scale, f"pair vs per-level kernels, tensor {j}"


def test_host_pipeline_real_chunk_size():
    """fwt._analysis_host_pipeline with its real 256 MB chunks (2.25 GB of input: 9 chunks of 64 images) against the
    device-resident path on the same data, and a mid-size input (cut into ~8 smaller chunks)."""
    from pytorch_wavelet_toolbox_b200 import fwt as F

    g = torch.Generator().manual_seed(77)
    for n in (9 * F.HOST_PIPELINE_CHUNK_BYTES // (1024 * 1024 * 4), 100):
        x = torch.randn(n, 1024, 1024, generator=g)
        assert x.numel() * 4 >= F.HOST_PIPELINE_MIN_BYTES
        host = wt.wavedec2(x, "db4", level=3)
        dev = wt.wavedec2(x.to(DEV), "db4", level=3)
        for a, b in zip(flatten_coeffs(host), flatten_coeffs(dev)):
            assert a.device.type == "cpu" and torch.equal(a, b.cpu())
        for i in (0, n // 2, n - 1):
            want = flatten_coeffs(P.wavedec2(x[i:i + 1], "db4", level=3))
            scale = max(float(t.abs().max()) for t in want)
            for a, b in zip(flatten_coeffs(host), want):
                assert_close_rel(a[i:i + 1], b, scale=scale, what=f"host pipeline image {i}")
        del host, dev, x
