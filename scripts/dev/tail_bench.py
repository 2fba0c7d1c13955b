"""Developer micro-benchmark of the fused decoder tail per site (not a pytest file): device time of the tail launch alone
from the library's CUDA-event profiler, one-tile-per-CTA kernel (option tail_persist = 0) next to the persistent
pipelined kernel (tail_persist = 1), at B = 1 and at a batch.  Bytes = SURVEY 8d's algorithmic traffic (fp32 element size)."""
import math
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import gpu_util as G  # noqa: E402
from oracle import synth  # noqa: E402

SITES = [
    # name, kind, C, S, groups, act, head couts, biases, batch for the throughput run
    ('upscaler02', 0, 32, 512, 32, 2, [7], [True], 8),
    ('morpher00', 0, 64, 256, 32, 2, [7], [True], 16),
    ('face_morpher08', 3, 64, 192, 0, 1, [2, 4, 1, 4, 1], [False, True, True, True, True], 32),
    ('combiner00', 2, 64, 128, 0, 1, [2, 1, 4, 1], [False, True, True, True], 32),
]
OUT_CH = {0: 15, 1: 18, 2: 24, 3: 24}


def run(site, N, persist, reps=5):
    name, kind, C, S, groups, act, couts, has_b, _ = site
    c = G.ctx()
    c.set_option('tail_persist', persist)
    g = torch.Generator().manual_seed(1)
    feature = torch.randn(1, C, S, S, generator=g).expand(N, C, S, S).contiguous() * 1.5 + 0.3
    gamma, beta = 1.0 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    ws = [torch.randn(co, C, 3, 3, generator=g) / math.sqrt(9 * C) * 0.1 for co in couts]
    bs = [0.1 * torch.randn(co, generator=g) if hb else None for co, hb in zip(couts, has_b)]
    image0 = synth.synthetic_image(3, 1)[:, :, :S, :S].expand(N, 4, S, S).contiguous()
    image1 = synth.synthetic_image(9, 1)[:, :, :S, :S].expand(N, 4, S, S).contiguous() if kind == 2 else None
    outs = G.tail(kind, feature, gamma, beta, groups, act, ws, bs, image0, image1, strict=0)
    c.set_option('profile', 2)
    for _ in range(reps):
        G.tail(kind, feature, gamma, beta, groups, act, ws, bs, image0, image1, strict=0)
    us = c.counter('prof_us_tail') / max(1, c.counter('prof_launches_tail'))
    c.set_option('profile', 0)
    img_ch = 8 if kind == 2 else 4
    mb = N * S * S * (C + img_ch + OUT_CH[kind]) * 4 / 1e6
    print('%-15s N %2d persist %d : %8.1f us  %7.1f MB  %7.1f GB/s  frac %.3f' % (name, N, persist, us, mb, mb / us * 1e3, mb / us * 1e3 / 6487.4), flush=True)
    return outs


if __name__ == '__main__':
    for site in SITES:
        for N in (1, site[-1]):
            a = run(site, N, 0)
            b = run(site, N, 1)
            d = max((x - y).abs().max().item() for x, y in zip(a, b))
            print('   max |one-tile - persistent| = %.3e' % d, flush=True)
    G.ctx().set_option('tail_persist', 1)
