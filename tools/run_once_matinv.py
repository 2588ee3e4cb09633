"""Two calls of MatrixWaverec on BASELINE configs[3] (1024 x 65536 f64 db6), for ncu."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt
x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
co = wt.MatrixWavedec("db6")(x)
iv = wt.MatrixWaverec("db6")
for _ in range(2):
    y = iv(co)
torch.cuda.synchronize()
print((y - x).abs().max().item())
