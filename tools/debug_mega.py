import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
from oracle import ptwt_port as P
torch.manual_seed(0)
for shape, mode, lev in (((4, 256, 256), 'zero', 2), ((5, 300, 200), 'reflect', 3), ((3, 1000, 520), 'symmetric', 4), ((8, 64, 64), 'periodic', 2),
                         ((23, 300, 420), 'reflect', 4), ((40, 96, 128), 'symmetric', 3)):
    x = torch.randn(*shape)
    c = wt.wavedec2(x.cuda(), 'db4', mode=mode, level=lev)
    torch.cuda.synchronize()
    w = P.wavedec2(x, 'db4', mode=mode, level=lev)
    fc = [c[0]] + [b for lv in c[1:] for b in lv]; fw = [w[0]] + [b for lv in w[1:] for b in lv]
    print(shape, mode, lev, 'max err %.3e' % max(float((a.cpu() - b).abs().max()) for a, b in zip(fc, fw)))
