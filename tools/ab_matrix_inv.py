"""MatrixWaverec float64: the DMMA synthesis cascade (matrix_dmma.cuh) against the per-level kernels (knob NO_DMMA) --
agreement over wavelets / lengths / level counts / odd lengths, then CUDA-graph replay times on BASELINE configs[3]
(1024 x 65536 f64 db6) for CTA size and chunk."""
import itertools, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt
from pytorch_wavelet_toolbox_b200 import _native

torch.manual_seed(0)
bad = 0
SWEEP = "--no-sweep" not in sys.argv
for wav, n, lev, bs in itertools.product(*[()] * 4) if not SWEEP else itertools.product(("haar", "db2", "db3", "db4", "sym5", "db6", "db7", "db8"),
                                         (64, 96, 250, 1000, 4096, 5001, 20000, 65536), (1, 2, 3, 5, None), (1, 7)):
    x = torch.randn(bs, n, device="cuda", dtype=torch.float64)
    try:
        fw = wt.MatrixWavedec(wav, level=lev)
        co = fw(x)
    except Exception as e:  # too many levels for this length etc.
        continue
    with _native.knobs(DISABLE_FUSED=1):
        co_ref = wt.MatrixWavedec(wav, level=lev)(x)
    with _native.knobs(MATF_VARIANT=1):
        co_v1 = wt.MatrixWavedec(wav, level=lev)(x)
    ferr = max(max((a - b).abs().max().item(), (c - b).abs().max().item()) for a, b, c in zip(co, co_ref, co_v1))
    if not ferr < 1e-12:
        bad += 1
        print(f"FORWARD MISMATCH {wav} n={n} level={lev} batch={bs}: {ferr:.3e}", flush=True)
    iv = wt.MatrixWaverec(wav)
    y = iv(co)
    with _native.knobs(NO_DMMA=1):
        iv2 = wt.MatrixWaverec(wav)
        y2 = iv2(co)
    with _native.knobs(MATI_ROWS=0):
        y3 = wt.MatrixWaverec(wav)(co)
    err = max((y - y2).abs().max().item(), (y3 - y2).abs().max().item())
    rt = (y[..., :n] - x).abs().max().item()
    ok = err < 1e-12 and rt < 1e-8
    if not ok:
        bad += 1
        print(f"MISMATCH {wav} n={n} level={lev} batch={bs}: vs per-level {err:.3e}, round trip {rt:.3e}", flush=True)
print("agreement sweep done, mismatches:", bad, flush=True)

x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
co = wt.MatrixWavedec("db6")(x)
iv = wt.MatrixWaverec("db6")
res = {}


def timeit(tag, **kn):
    with _native.knobs(**kn):
        for _ in range(3):
            iv(co)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = iv(co)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        del g, keep
    res[tag] = ms
    print(f"{tag}: {ms:.3f} ms = {2 * x.numel() * 8 / ms / 1e6 / 6501.9 * 100:.1f} %", flush=True)


timeit("per-level", NO_DMMA=1)
timeit("dmma default")

# the analysis side
fw = wt.MatrixWavedec("db6")
iv = fw
co = x
timeit("forward default (polyphase kernel)")
timeit("forward streaming kernel", MATF_VARIANT=1)
for chunk, kc, nt in itertools.product((1024, 2048, 4096), (2, 4, 8), (128, 256)):
    timeit(f"forward polyphase chunk{chunk} kcoarse{kc} nt{nt}", MATF_CHUNK=chunk, MATF_KCOARSE=kc, MATF_NT=nt)
timeit("forward polyphase k=3", MATF_K=3)
timeit("forward polyphase k=4", MATF_K=4)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/ab_matrix_inv.json").write_text(json.dumps(res, indent=1))
