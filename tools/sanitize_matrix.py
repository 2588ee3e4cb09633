"""Small float64 matrix-FWT cases through every DMMA kernel variant, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_matrix.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt
from pytorch_wavelet_toolbox_b200 import _native

torch.manual_seed(0)
worst = 0.0
for wav, n, lev, bs in (("haar", 64, 3, 3), ("db2", 250, None, 2), ("db4", 1000, None, 3), ("db6", 5001, None, 2),
                        ("db8", 8192, None, 2), ("db6", 20000, 4, 3), ("sym5", 4096, 5, 5)):
    x = torch.randn(bs, n, device="cuda", dtype=torch.float64)
    for kn in ({}, {"MATF_VARIANT": 1}, {"MATI_ROWS": -2}, {"MATI_ROWS": 4}, {"MATF_NT": 256, "MATI_NT": 256},
               {"MATF_K": 4, "MATI_K": 4}):
        with _native.knobs(**kn):
            co = wt.MatrixWavedec(wav, level=lev)(x)
            y = wt.MatrixWaverec(wav)(co)
        worst = max(worst, (y[..., :n] - x).abs().max().item())
torch.cuda.synchronize()
print("worst round trip", worst)
assert worst < 1e-9
