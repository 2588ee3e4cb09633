"""Achieved algorithmic GB/s of the other BASELINE.json configs (parity-test cases, not bench lines)."""
import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt


def timeit(fn, n=10):
    """(ms per call enqueued eagerly, result, host ms per call to enqueue, ms per replay of the call captured in a CUDA graph)."""
    import time
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / n
    graph_ms = float("nan")
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / n
        del keep, g
    except Exception as ex:  # noqa: BLE001
        print("   (graph capture failed:", type(ex).__name__, str(ex)[:100], ")")
    return eager, r, host, graph_ms


def nbytes(t):
    if isinstance(t, torch.Tensor):
        return t.numel() * t.element_size()
    if isinstance(t, dict):
        return sum(nbytes(v) for v in t.values())
    return sum(nbytes(v) for v in t)


def report(name, t, alg):
    ms, host, graph = t
    best = min(ms, graph) if graph == graph else ms
    print(f"{name:58s} eager {ms:7.3f} ms (host enqueue {host:6.3f} ms)  graph replay {graph:7.3f} ms  -> "
          f"{alg / best / 1e6:7.1f} GB/s algorithmic = {alg / best / 1e6 / 6501.9 * 100:5.1f}% of 6501.9")


which = sys.argv[1:] or ["1d", "3d", "mat", "db8"]
if "1d" in which:
    x = torch.randn(32, 1_000_000, device="cuda")
    ms, c, host, gr = timeit(lambda: wt.wavedec(x, "db5", mode="periodic", level=10))
    report("wavedec db5 L10 periodic 32x1e6 f32 (speed-test cfg)", (ms, host, gr), nbytes(x) + nbytes(c))
    ms, r, host, gr = timeit(lambda: wt.waverec(c, "db5"))
    report("waverec of it", (ms, host, gr), nbytes(r) + nbytes(c))
    ms, c, host, gr = timeit(lambda: wt.wavedec(x, "db5", mode="reflect", level=10))
    report("wavedec db5 L10 reflect 32x1e6 f32 (fused multi-level kernel)", (ms, host, gr), nbytes(x) + nbytes(c))
if "3d" in which:
    x = torch.randn(8, 256, 256, 256, device="cuda")
    ms, c, host, gr = timeit(lambda: wt.wavedec3(x, "sym4", level=3))
    report("wavedec3 sym4 L3 zero 8x256^3 f32 (BASELINE cfg 3)", (ms, host, gr), nbytes(x) + nbytes(c))
    ms, r, host, gr = timeit(lambda: wt.waverec3(c, "sym4"))
    report("waverec3 of it", (ms, host, gr), nbytes(r) + nbytes(c))
if "mat" in which:
    x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
    fw = wt.MatrixWavedec("db6")
    ms, c, host, gr = timeit(lambda: fw(x))
    report("MatrixWavedec db6 1024x65536 f64 L12 (BASELINE cfg 4)", (ms, host, gr), nbytes(x) + nbytes(c))
    inv = wt.MatrixWaverec("db6")
    ms, r, host, gr = timeit(lambda: inv(c))
    report("MatrixWaverec of it", (ms, host, gr), nbytes(r) + nbytes(c))
if "mat2" in which:
    x = torch.randn(32, 2048, 2048, device="cuda")
    fw2 = wt.MatrixWavedec2("db4", 4)
    ms, c, host, gr = timeit(lambda: fw2(x))
    report("MatrixWavedec2 db4 L4 32x2048^2 f32 (separable, 8f row 2)", (ms, host, gr), 2 * nbytes(x))
    inv2 = wt.MatrixWaverec2("db4")
    ms, r, host, gr = timeit(lambda: inv2(c))
    report("MatrixWaverec2 of it", (ms, host, gr), 2 * nbytes(x))
if "db8" in which:
    x = torch.randn(128, 2048, 2048, device="cuda")
    ms, c, host, gr = timeit(lambda: wt.wavedec2(x, "db8", level=5))
    report("wavedec2 db8 L5 128x2048^2 f32 (BASELINE cfg 5 per-GPU shape, 1/4 batch)", (ms, host, gr), nbytes(x) + nbytes(c))
    ms, r, host, gr = timeit(lambda: wt.waverec2(c, "db8"))
    report("waverec2 of it", (ms, host, gr), nbytes(r) + nbytes(c))
    x = torch.randn(16, 4096, 4096, device="cuda", dtype=torch.float64)
    ms, c, host, gr = timeit(lambda: wt.wavedec2(x, "db4", level=4))
    report("wavedec2 db4 L4 16x4096^2 f64", (ms, host, gr), nbytes(x) + nbytes(c))
