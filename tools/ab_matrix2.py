"""Sweep of the fused matrix-FWT analysis kernel on BASELINE configs[3] (1024 x 65536 f64 db6): CTA size, chunk,
levels per launch, register cap.  CUDA-graph replay times (ms) so that host cost does not blur the kernel."""
import itertools, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt
from pytorch_wavelet_toolbox_b200 import _native

x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
fw = wt.MatrixWavedec("db6")
fw(x)
res = {}
for nt, chunk, k, minb, cpc in itertools.product((128, 256), (1024, 2048, 4096), (4,), (1, 2), (1, 4, 8, 16)):
    if nt == 128 and chunk > 4096:
        continue
    with _native.knobs(MATF_NT=nt, MATF_CHUNK=chunk, MATF_MINB=minb, MATF_CPC=cpc):
        for _ in range(3):
            fw(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fw(x)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        del g, keep
    res[f"nt{nt} chunk{chunk} k{k} minb{minb} cpc{cpc}"] = ms
    print(f"nt {nt:3d} chunk {chunk:5d} cpc {cpc:2d} minb {minb}: {ms:.3f} ms = {2 * x.numel() * 8 / ms / 1e6 / 6501.9 * 100:.1f} %", flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/ab_matrix2.json").write_text(json.dumps(res, indent=1))
