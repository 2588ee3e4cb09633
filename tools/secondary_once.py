"""One call of each non-headline configuration (for `ncu -k regex:...` captures; see profiles/README.md)."""
import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt

which = sys.argv[1:] or ["3d", "mat", "1d"]
if "3d" in which:
    x = torch.randn(8, 256, 256, 256, device="cuda")
    c = wt.wavedec3(x, "sym4", level=3)
    wt.waverec3(c, "sym4")
if "mat" in which:
    x = torch.randn(1024, 65536, device="cuda", dtype=torch.float64)
    c = wt.MatrixWavedec("db6")(x)
    wt.MatrixWaverec("db6")(c)
if "1d" in which:
    x = torch.randn(32, 1_000_000, device="cuda")
    c = wt.wavedec(x, "db5", mode="periodic", level=10)
    wt.waverec(c, "db5")
torch.cuda.synchronize()
