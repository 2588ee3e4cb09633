"""A/B of the fused matrix-transform tunables on BASELINE config 4 (graph replay = device-bound time)."""
import os, sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt

x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
fw, iv = wt.MatrixWavedec("db6"), wt.MatrixWaverec("db6")
c = fw(x)
ref = iv(c)
alg = 2 * x.numel() * 8


def graph_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, keep


def setenv(d):
    for k in ("WTB200_MATF_K", "WTB200_MATF_CHUNK", "WTB200_MATI_K", "WTB200_MATI_CHUNK"):
        os.environ.pop(k, None)
    os.environ.update(d)


for k, ch in ((1, 4096), (6, 4096), (4, 4096), (4, 8192), (5, 8192), (3, 4096), (4, 6144)):
    setenv({"WTB200_MATF_K": str(k), "WTB200_MATF_CHUNK": str(ch)})
    ms, got = graph_ms(lambda: fw(x))
    err = max(float((a - b).abs().max()) for a, b in zip(got, c))
    print(f"fwd K={k} chunk={ch:5d}: {ms:.3f} ms  {alg / ms / 1e6 / 6501.9 * 100:5.1f}%  max diff vs default {err:.1e}")
for k, ch in ((1, 2048), (2, 4096), (3, 4096), (4, 8192)):
    setenv({"WTB200_MATI_K": str(k), "WTB200_MATI_CHUNK": str(ch)})
    ms, got = graph_ms(lambda: iv(c))
    print(f"inv K={k} chunk={ch:5d}: {ms:.3f} ms  {alg / ms / 1e6 / 6501.9 * 100:5.1f}%  round trip {float((got - x).abs().max()):.1e}")
