"""Developer probe: per-phase clock64 stamps of the halo conv kernel (THA4_HALO_DEBUG=1) on the big teacher layer shapes."""
import os
os.environ['THA4_HALO_DEBUG'] = '1'
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'tests'))
import torch
import gpu_util as G

torch.manual_seed(0)
for (C, Co, S) in [(32, 32, 512), (64, 64, 256), (128, 128, 128), (32, 64, 512), (256, 256, 64), (512, 512, 32)]:
    x = torch.randn(1, C, S, S)
    w = torch.randn(Co, C, 3, 3) * 0.05
    g = torch.ones(C); b = torch.zeros(C)
    print(f'--- C={C} Cout={Co} S={S}', file=sys.stderr, flush=True)
    G.conv_norm(0, x, C, C, g, b, None, None, 1, w)
