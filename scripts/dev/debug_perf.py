"""Developer micro-benchmark of single conv layers (not a pytest file): device time of the conv launch sequence
(conv + optional memset / reduce kernel) from the library's own CUDA-event profiler."""
import math
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_util as G  # noqa: E402


def bench(kind, N, Cin, H, Cout, ksplit=0, cluster=1, reps=20, persistent=0, mt2=0, half=0):
    c = G.ctx()
    c.set_option('half_operands', half)
    c.set_option('cluster_splitk', cluster)
    c.set_option('persistent_conv', persistent)
    c.set_option('conv_mt2', mt2)
    g = torch.Generator().manual_seed(0)
    k = {0: 3, 3: 1}[kind]
    x = torch.randn(N, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    G.conv(kind, x, w, None, None, 0, 0, strict=0, ksplit=ksplit)
    c.set_option('profile', 2)
    for _ in range(reps):
        G.conv(kind, x, w, None, None, 0, 0, strict=0, ksplit=ksplit)
    us = c.counter('prof_us_conv') / reps
    c.set_option('profile', 0)
    c.set_option('cluster_splitk', 1)
    gf = 2.0 * N * H * H * Cout * Cin * k * k / 1e9
    c.set_option('persistent_conv', 0)
    c.set_option('conv_mt2', 0)
    c.set_option('half_operands', 1)
    print('kind %d N%d Cin%4d H%3d Cout%4d ksplit %2d cluster %d persistent %d mt2 %d half %d : %7.1f us  %7.1f TFLOP/s' % (kind, N, Cin, H, Cout, ksplit, cluster, persistent, mt2, half, us, gf / us * 1e3))


if __name__ == '__main__':
    for shape in [(0, 1, 256, 16, 256), (0, 1, 512, 16, 512), (0, 1, 256, 32, 256), (0, 1, 256, 64, 256), (0, 1, 128, 128, 128),
                  (0, 1, 64, 256, 64), (0, 1, 32, 512, 32), (0, 1, 64, 512, 64), (0, 1, 128, 256, 128), (0, 1, 256, 128, 256)]:
        for ks, cl in [(0, 1), (0, 0), (1, 0), (2, 1), (4, 1), (8, 1), (4, 0), (8, 0), (16, 0)]:
            if shape[3] >= 128 and ks not in (0, 1):
                continue
            bench(*shape, ksplit=ks, cluster=cl)
        print()
