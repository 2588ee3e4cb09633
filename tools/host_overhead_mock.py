"""Host-side cost of one call with the native library and CUDA mocked out (runs without a GPU).
Only for profiling the Python layer; nothing is computed."""
import contextlib, cProfile, pstats, sys, time, types
import torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
from pytorch_wavelet_toolbox_b200 import fwt, matrix_fwt, _native as N


class _Lib:
    def __getattr__(self, name):
        if name == "wt_dwt_workspace_bytes":
            return lambda *a: 0
        return lambda *a: 0


class _Stream:
    cuda_stream = 0
    def synchronize(self): pass


cpu = torch.device("cpu")
for mod in (fwt, matrix_fwt):
    if hasattr(mod, "_compute_device"):
        mod._compute_device = lambda t: cpu
N.load = lambda: _Lib()
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.cuda.current_stream = lambda d=None: _Stream()
_is_cuda = property(lambda self: True)
torch.Tensor.is_cuda = _is_cuda


def bench(name, fn, n=200):
    fn(); fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t) / n
    print(f"{name:40s} {dt * 1e6:8.1f} us/call")
    if "-p" in sys.argv:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(n): fn()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


x1 = torch.randn(4, 100_000)
c1 = wt.wavedec(x1, "db5", mode="periodic", level=10)
bench("wavedec 1-D L10", lambda: wt.wavedec(x1, "db5", mode="periodic", level=10))
bench("waverec 1-D L10", lambda: wt.waverec(c1, "db5"))
x2 = torch.randn(2, 512, 512)
c2 = wt.wavedec2(x2, "db4", level=4)
bench("wavedec2 L4", lambda: wt.wavedec2(x2, "db4", level=4))
bench("waverec2 L4", lambda: wt.waverec2(c2, "db4"))
x3 = torch.randn(2, 64, 64, 64)
c3 = wt.wavedec3(x3, "sym4", level=3)
bench("wavedec3 L3", lambda: wt.wavedec3(x3, "sym4", level=3))
bench("waverec3 L3", lambda: wt.waverec3(c3, "sym4"))
xm = torch.randn(8, 65536, dtype=torch.float64)
fw = wt.MatrixWavedec("db6"); cm = fw(xm)
bench("MatrixWavedec L12", lambda: fw(xm))
iv = wt.MatrixWaverec("db6"); iv(cm)
bench("MatrixWaverec L12", lambda: iv(cm))
