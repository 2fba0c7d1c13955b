"""ncu target: the streaming (HBM / L2 bound) conv layers in isolation (developer script)."""
import os, sys, math, torch
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G
g = torch.Generator().manual_seed(0)
def one(N, C, H, Co, res):
    x = torch.randn(N, C, H, H, generator=g); w = torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(C * 9)
    b = torch.randn(Co, generator=g)
    r = torch.randn(N, Co, H, H, generator=g) if res else None
    for _ in range(2):
        G.conv(0, x, w, b, r, 1 if res else 0, 0, strict=0, ksplit=0)
one(8, 32, 512, 32, True)
one(1, 32, 512, 32, True)
one(8, 64, 256, 64, True)
