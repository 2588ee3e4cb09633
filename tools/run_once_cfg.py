"""One call of a BASELINE configuration (for ncu): python tools/run_once_cfg.py {2,3,4,5} [batch] [level]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import pytorch_wavelet_toolbox_b200 as wt
cfg = bench.CONFIGS[int(sys.argv[1])]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["shape"][0]
lev = int(sys.argv[3]) if len(sys.argv) > 3 else "cfg"
x = torch.randn((B,) + tuple(cfg["shape"][1:]), device="cuda", dtype=bench.DT[cfg["dtype"]])
f = bench.make_forward(wt, cfg, level=lev)
for _ in range(2):
    c = f(x)
torch.cuda.synchronize()
