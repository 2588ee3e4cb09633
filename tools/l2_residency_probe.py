"""Does the level-1 approximation survive in L2 until the level-2 kernel reads it?  Times the 4-level
transform per image for small sub-batches (approximation bands fit in the 126 MB L2) vs the full batch."""
import sys, torch
sys.path.insert(0, '.')
import pytorch_wavelet_toolbox_b200 as wt
x = torch.randn(64, 4096, 4096, device='cuda')
for bs in (64, 16, 8, 4, 2, 1):
    for _ in range(3):
        for i in range(0, 64, bs):
            wt.wavedec2(x[i:i + bs], 'db4', level=4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for i in range(0, 64, bs):
            wt.wavedec2(x[i:i + bs], 'db4', level=4)
    e1.record(); torch.cuda.synchronize()
    print(f"sub-batch {bs:3d}: {e0.elapsed_time(e1) / 5:.3f} ms per 64 images")
