"""Seeded synthetic weights and inputs for the THA4 hot path  --  TEST INFRASTRUCTURE.

The teacher weights are not shipped with the reference and 657 MB of fp32 cannot be
committed, so every test/bench regenerates them from a seed with torch's CPU generator (bit-identical
across machines running the same torch build).  The key/shape lists below restate the reference's
state_dict layout (verified key-for-key against the live reference modules by
tests/test_oracle_pinned.py when /root/reference is present).

Tensors that the reference zero-initialises (U-Net conv1 / last / attention out-proj,
coarse_image_conv, grid_change heads) are drawn from small normals instead, otherwise a random-init teacher
outputs zero warps and the parity tests would exercise nothing.
"""
import math
from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]   # (key, shape, role)


def _enc_dec_spec(prefix: str, in_ch: int, pose_ch: int) -> Spec:
    """Keys of PoserEncoderDecoder00 / FaceMorpher08 trunk (poser_encoder_decoder_00.py:50-91)."""
    s: Spec = []
    chans = [64, 128, 256, 512]
    s += [(prefix + 'downsample_blocks.0.0.weight', (64, in_ch, 3, 3), 'conv'),
          (prefix + 'downsample_blocks.0.1.weight', (64,), 'norm_w'),
          (prefix + 'downsample_blocks.0.1.bias', (64,), 'norm_b')]
    for i in range(1, 4):
        s += [(prefix + 'downsample_blocks.%d.0.weight' % i, (chans[i], chans[i - 1], 4, 4), 'conv'),
              (prefix + 'downsample_blocks.%d.1.weight' % i, (chans[i],), 'norm_w'),
              (prefix + 'downsample_blocks.%d.1.bias' % i, (chans[i],), 'norm_b')]
    s += [(prefix + 'bottleneck_blocks.0.0.weight', (512, 512 + pose_ch, 3, 3), 'conv'),
          (prefix + 'bottleneck_blocks.0.1.weight', (512,), 'norm_w'),
          (prefix + 'bottleneck_blocks.0.1.bias', (512,), 'norm_b')]
    for i in range(1, 6):
        for j, k in ((0, 1), (3, 4)):
            s += [(prefix + 'bottleneck_blocks.%d.resnet_path.%d.weight' % (i, j), (512, 512, 3, 3), 'conv'),
                  (prefix + 'bottleneck_blocks.%d.resnet_path.%d.weight' % (i, k), (512,), 'norm_w'),
                  (prefix + 'bottleneck_blocks.%d.resnet_path.%d.bias' % (i, k), (512,), 'norm_b')]
    for i in range(3):
        cin, cout = chans[3 - i], chans[2 - i]
        s += [(prefix + 'upsample_blocks.%d.0.weight' % i, (cin, cout, 4, 4), 'convT'),
              (prefix + 'upsample_blocks.%d.1.weight' % i, (cout,), 'norm_w'),
              (prefix + 'upsample_blocks.%d.1.bias' % i, (cout,), 'norm_b')]
    return s


def _head_spec(name: str, cout: int, bias: bool, role: str = 'conv') -> Spec:
    s: Spec = [(name + '.weight', (cout, 64, 3, 3), role)]
    if bias:
        s.append((name + '.bias', (cout,), 'bias'))
    return s


def eyebrow_decomposer_spec() -> Spec:
    s = _enc_dec_spec('body.', 4, 0)
    for name, c in (('background_layer_alpha.0', 1), ('background_layer_color_change.0', 4),
                    ('eyebrow_layer_alpha.0', 1), ('eyebrow_layer_color_change.0', 4)):
        s += _head_spec(name, c, True)
    return s


def eyebrow_morphing_combiner_spec() -> Spec:
    s = _enc_dec_spec('body.', 8, 12)
    s += _head_spec('morphed_eyebrow_layer_grid_change', 2, False, 'grid_head')
    s += _head_spec('morphed_eyebrow_layer_alpha.0', 1, True)
    s += _head_spec('morphed_eyebrow_layer_color_change.0', 4, True)
    s += _head_spec('combine_alpha.0', 1, True)
    return s


def face_morpher_spec() -> Spec:
    s = _enc_dec_spec('', 4, 27)
    s += _head_spec('iris_mouth_grid_change', 2, False, 'grid_head')
    s += _head_spec('iris_mouth_color_change.0', 4, True)
    s += _head_spec('iris_mouth_alpha.0', 1, True)
    s += _head_spec('eye_color_change.0', 4, True)
    s += _head_spec('eye_alpha.0', 1, True)
    return s


def _res_block_spec(p: str, cin: int, cout: int) -> Spec:
    """Registration order of ResBlock.__init__ (unet.py:118-152)."""
    s: Spec = [(p + '.norm0.weight', (cin,), 'norm_w'), (p + '.norm0.bias', (cin,), 'norm_b'),
               (p + '.conv0.weight', (cout, cin, 3, 3), 'conv'), (p + '.conv0.bias', (cout,), 'bias'),
               (p + '.cond0_layers.1.weight', (2 * cout, 256), 'film'), (p + '.cond0_layers.1.bias', (2 * cout,), 'bias'),
               (p + '.norm1.weight', (cout,), 'norm_w'), (p + '.norm1.bias', (cout,), 'norm_b'),
               (p + '.conv1.weight', (cout, cout, 3, 3), 'zconv'), (p + '.conv1.bias', (cout,), 'bias'),
               (p + '.cond1_layers.1.weight', (2 * cout, 256), 'film'), (p + '.cond1_layers.1.bias', (2 * cout,), 'bias')]
    if cin != cout:
        s += [(p + '.skip.weight', (cout, cin, 1, 1), 'conv'), (p + '.skip.bias', (cout,), 'bias')]
    return s


def _attn_spec(p: str, c: int) -> Spec:
    return [(p + '.norm.weight', (c,), 'norm_w'), (p + '.norm.bias', (c,), 'norm_b'),
            (p + '.qkv.weight', (3 * c, c, 1, 1), 'conv1'), (p + '.qkv.bias', (3 * c,), 'bias'),
            (p + '.conv.weight', (c, c, 1, 1), 'zconv'), (p + '.conv.bias', (c,), 'bias')]


def unet_spec(p: str, model_channels: int, mults: List[int], use_attention: List[bool]) -> Spec:
    """Keys of Unet / UnetWithFirstConvAddition (unet.py:438-529) for 1 res-block per level, 4 middle res-blocks."""
    mc = model_channels
    s: Spec = [(p + 'time_embed.1.weight', (256, mc), 'linear'), (p + 'time_embed.1.bias', (256,), 'bias'),
               (p + 'time_embed.3.weight', (256, 256), 'linear'), (p + 'time_embed.3.bias', (256,), 'bias'),
               (p + 'cond_embed.0.weight', (256, 6), 'linear'), (p + 'cond_embed.0.bias', (256,), 'bias'),
               (p + 'cond_embed.2.weight', (256, 256), 'linear'), (p + 'cond_embed.2.bias', (256,), 'bias'),
               (p + 'first_conv.weight', (mc, 4, 3, 3), 'conv'), (p + 'first_conv.bias', (mc,), 'bias')]
    cur = mc
    channels = [cur]
    L = len(mults)
    for i in range(L):
        out = mc * mults[i]
        bp = p + 'down_blocks.%d' % i
        s += _res_block_spec(bp + '.res_blocks.0', cur, out)
        if use_attention[i]:
            s += _attn_spec(bp + '.attention_blocks.0', out)
        channels.append(out)
        if i < L - 1:
            s += _res_block_spec(bp + '.downsample', out, out)
            channels.append(out)
        cur = out
    for j in range(7):
        mp = p + 'middle_blocks.%d' % j
        s += _res_block_spec(mp, cur, cur) if j % 2 == 0 else _attn_spec(mp + '.module', cur)
    for bi, i in enumerate(reversed(range(L))):
        out = mc * mults[i]
        bp = p + 'up_blocks.%d' % bi
        blocks: Spec = []
        attn: Spec = []
        for r in range(2):
            skip = channels.pop()
            blocks += _res_block_spec(bp + '.resnet_blocks.%d' % r, (cur if r == 0 else out) + skip, out)
            if use_attention[i]:
                attn += _attn_spec(bp + '.attention_blocks.%d' % r, out)
        s += blocks + attn
        if i > 0:
            s += _res_block_spec(bp + '.upsample', out, out)
        cur = out
    assert not channels
    s += [(p + 'last.0.weight', (cur,), 'norm_w'), (p + 'last.0.bias', (cur,), 'norm_b'),
          (p + 'last.2.weight', (7, cur, 3, 3), 'last'), (p + 'last.2.bias', (7,), 'bias')]
    return s


def body_morpher_spec() -> Spec:
    return unet_spec('body.', 64, [1, 2, 4, 4, 4], [False] * 4 + [True])


def upscaler_spec() -> Spec:
    return unet_spec('body.', 32, [1, 2, 4, 8, 8, 8], [False] * 5 + [True]) + [
        ('coarse_image_conv.weight', (32, 10, 3, 3), 'zconv'), ('coarse_image_conv.bias', (32,), 'bias')]


def siren_face_morpher_spec() -> Spec:
    s: Spec = []
    cin = 41
    for i in range(8):
        s += [('siren.sine_layers.%d.linear.weight' % i, (128, cin, 1, 1), 'siren_first' if i == 0 else 'siren'),
              ('siren.sine_layers.%d.linear.bias' % i, (128,), 'siren_bias')]
        cin = 128
    s += [('siren.last_linear.weight', (4, 128, 1, 1), 'conv1'), ('siren.last_linear.bias', (4,), 'bias')]
    return s


def siren_morpher_03_spec() -> Spec:
    s: Spec = []
    dims = [[(47, 360), (360, 360), (360, 180)], [(227, 180), (180, 180), (180, 90)], [(137, 90), (90, 90), (90, 90)]]
    for i, lv in enumerate(dims):
        for j, (ci, co) in enumerate(lv):
            s += [('siren_layers.%d.%d.linear.weight' % (i, j), (co, ci, 1, 1), 'siren_first' if (i, j) == (0, 0) else 'siren'),
                  ('siren_layers.%d.%d.linear.bias' % (i, j), (co,), 'siren_bias')]
    s += [('last_linear.weight', (7, 90, 1, 1), 'student_last'), ('last_linear.bias', (7,), 'bias')]
    return s


TEACHER_SPECS = {
    'eyebrow_decomposer': eyebrow_decomposer_spec,
    'eyebrow_morphing_combiner': eyebrow_morphing_combiner_spec,
    'face_morpher': face_morpher_spec,
    'body_morpher': body_morpher_spec,
    'upscaler': upscaler_spec,
}
STUDENT_SPECS = {'face_morpher': siren_face_morpher_spec, 'body_morpher': siren_morpher_03_spec}


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
