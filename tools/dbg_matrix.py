import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
for dt in (torch.float64, torch.float32):
    for n, lev in ((4096, 6), (65536, 12), (256, 3)):
        x = torch.randn(4, n, device="cuda", dtype=dt)
        c = wt.MatrixWavedec("db6", lev)(x)
        torch.cuda.synchronize()
        print(dt, n, lev, "ok", float(c[0].abs().max()))
