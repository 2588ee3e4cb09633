import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
from oracle import ptwt_port as P
dt = torch.float32 if len(sys.argv) < 2 or sys.argv[1] == 'f32' else torch.float64
wav = sys.argv[2] if len(sys.argv) > 2 else 'db4'
mode = sys.argv[3] if len(sys.argv) > 3 else 'zero'
x = torch.randn(2, 64, 64, dtype=dt)
c = wt.wavedec2(x.cuda(), wav, mode=mode, level=1)
torch.cuda.synchronize()
w = P.wavedec2(x, wav, mode=mode, level=1)
print("err", max(float((a.cpu()-b).abs().max()) for a, b in zip([c[0], *c[1]], [w[0], *w[1]])))
