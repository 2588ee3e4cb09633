"""A/B of the 2-D analysis paths on BASELINE configs[1] (64 x 4096^2 f32 db4 L4): the two-level kernel of independent
warps (default) against one launch per level, plus a sweep of its segment length.  CUDA events, 20 iterations."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt  # noqa: E402
from pytorch_wavelet_toolbox_b200 import _native  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    wav = sys.argv[2] if len(sys.argv) > 2 else "db4"
    x = torch.randn(B, 4096, 4096, device="cuda")
    alg = 134612360 * B
    res = {}
    for name, kn, lev in (
        ("default L4", {}, 4), ("wpair L4", {"WPAIR": 1}, 4),
        ("hybrid 8", {"WPAIR_HYBRID": 8}, 4), ("hybrid 12", {"WPAIR_HYBRID": 12}, 4), ("hybrid 16", {"WPAIR_HYBRID": 16}, 4),
        ("hybrid 20", {"WPAIR_HYBRID": 20}, 4), ("hybrid 24", {"WPAIR_HYBRID": 24}, 4), ("hybrid 32", {"WPAIR_HYBRID": 32}, 4),
        ("default L4 again", {}, 4),
    ):
        with _native.knobs(**kn):
            _native.launch_count_reset()
            med, mn = timeit(lambda: wt.wavedec2(x, wav, level=lev))
            nl = _native.launch_count() // 25
        res[name] = {"median_ms": med, "min_ms": mn, "launches": nl}
        if lev == 4:
            res[name]["step_frac"] = alg / (med * 1e-3) / 1e9 / 6501.9
        print(name, res[name], flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/ab_wpair.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
