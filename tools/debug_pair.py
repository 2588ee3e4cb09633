import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
from oracle import ptwt_port as P
torch.manual_seed(0)
for shape, mode in (((1, 256, 256), 'zero'), ((1, 256, 256), 'reflect'), ((2, 300, 200), 'reflect')):
    x = torch.randn(*shape)
    c = wt.wavedec2(x.cuda(), 'db4', mode=mode, level=2)
    torch.cuda.synchronize()
    w = P.wavedec2(x, 'db4', mode=mode, level=2)
    names = ['cA2', 'H2', 'V2', 'D2', 'H1', 'V1', 'D1']
    flat_c = [c[0], *c[1], *c[2]]; flat_w = [w[0], *w[1], *w[2]]
    for n, a, b in zip(names, flat_c, flat_w):
        d = (a.cpu() - b).abs()
        bad = (d > 1e-4).nonzero()
        print(shape, mode, n, tuple(a.shape), 'max err %.3e' % d.max().item(), 'nbad', bad.shape[0],
              'first bad', bad[:3].tolist() if bad.numel() else '')
