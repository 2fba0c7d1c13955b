"""Seeded synthetic weights and inputs of the THA4 hot path (benchmarks, smoke tests, parity tests).

The teacher weights are not shipped with the reference (they are a separate download, README.md:164-183 there) and
657 MB of fp32 cannot be committed, so benchmarks and tests regenerate them from a seed with torch's CPU generator
(bit-identical across machines running the same torch build).  Tensors that the reference zero-initialises (U-Net
conv1 / last / attention out-proj, coarse_image_conv, grid_change heads) are drawn from small normals instead,
otherwise a random-init teacher outputs zero warps and nothing downstream is exercised.
"""
import math
from typing import Dict

import torch

from tha4_b200.nn.state_dict_spec import Spec, STUDENT_SPECS, TEACHER_SPECS


def _draw(shape, role: str, g: torch.Generator) -> torch.Tensor:
    z = torch.randn(shape, generator=g, dtype=torch.float32)
    if role in ('conv', 'conv1'):
        fan_in = shape[1] * shape[2] * shape[3]
        return z * math.sqrt(2.0 / fan_in)
    if role == 'convT':      # kaiming_normal_ on [Cin, Cout, k, k]: torch's fan_in = size(1) * k * k
        return z * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
    if role == 'zconv':      # reference zero-inits; keep the residual branch at ~half scale
        return z * (0.5 * math.sqrt(1.0 / (shape[1] * shape[2] * shape[3])))
    if role == 'grid_head':  # reference zero-inits; warp offsets of ~0.05 in normalised coordinates
        return z * 0.003
    if role == 'last':       # reference zero-inits; direct ~0.4, grid ~0.04, alpha logit ~0.4
        w = z * (0.03 * math.sqrt(64.0 / shape[1]))
        w[4:6] *= 0.1
        return w
    if role == 'student_last':
        w = z * math.sqrt(2.0 / shape[1])
        w[0:2] *= 0.05
        return w
    if role == 'linear':
        return z * math.sqrt(1.0 / shape[1])
    if role == 'film':
        return z * (0.5 / math.sqrt(shape[1]))
    if role == 'norm_w':
        return 1.0 + 0.1 * z
    if role == 'norm_b':
        return 0.1 * z
    if role == 'bias':
        return 0.05 * z
    if role == 'siren_first':   # uniform(-1/in, 1/in) in the reference (siren.py:32); same scale here
        return (torch.rand(shape, generator=g) * 2 - 1) / shape[1]
    if role == 'siren':         # uniform(+-sqrt(6/in)/30) (siren.py:34-36)
        return (torch.rand(shape, generator=g) * 2 - 1) * (math.sqrt(6.0 / shape[1]) / 30.0)
    if role == 'siren_bias':
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    raise ValueError(role)


def make_state_dict(spec: Spec, seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return {k: _draw(shape, role, g) for k, shape, role in spec}


def teacher_state_dicts(seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    return {name: make_state_dict(fn(), seed * 100 + i) for i, (name, fn) in enumerate(TEACHER_SPECS.items())}


def student_state_dicts(seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    return {name: make_state_dict(fn(), seed * 100 + 50 + i) for i, (name, fn) in enumerate(STUDENT_SPECS.items())}


def synthetic_image(seed: int = 0, n: int = 1) -> torch.Tensor:
    """[n,4,512,512] in [-1,1]: smooth premultiplied-alpha-like blobs (not white noise, so that warps and
    bilinear taps see realistic gradients)."""
    g = torch.Generator().manual_seed(1000 + seed)
    low = torch.rand(n, 4, 32, 32, generator=g)
    img = torch.nn.functional.interpolate(low, size=(512, 512), mode='bicubic', align_corners=False).clamp(0, 1)
    img = img + 0.05 * torch.rand(n, 4, 512, 512, generator=g)
    alpha = img[:, 3:4].clamp(0, 1)
    img = torch.cat([img[:, 0:3].clamp(0, 1) * alpha, alpha], dim=1)
    return (img * 2.0 - 1.0).contiguous()


def random_poses(n: int, seed: int = 1234) -> torch.Tensor:
    """pose_i ~ U(range_i): morph parameters [0,1]; iris_rotation_x/y, head_x/y, neck_z, body_y/z in [-1,1];
    breathing [0,1]  (poser/modes/pose_parameters.py:6-35)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, 45, generator=g)
    lo = torch.zeros(45)
    lo[37:44] = -1.0
    return lo + u * (1.0 - lo)
