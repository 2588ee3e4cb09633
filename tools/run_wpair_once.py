import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pytorch_wavelet_toolbox_b200 as wt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lev = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wav = sys.argv[3] if len(sys.argv) > 3 else "db4"
x = torch.randn(B, 4096, 4096, device="cuda")
for _ in range(2):
    c = wt.wavedec2(x, wav, level=lev)
torch.cuda.synchronize()
