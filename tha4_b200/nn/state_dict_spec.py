"""state_dict layout (key, shape, init role) of the seven THA4 networks, in the reference's registration order.

This is the on-disk weight format of the reference (`torch.save(module.state_dict())`,
src/tha4/shion/core/load_save.py:6-14); the modules in tha4_b200.nn register exactly these keys so the shipped
`.pt` files load unchanged.  Roles name the reference initialiser of each tensor (used by `reset_parameters`):
conv/conv1/convT = He normal (nn/init_function.py:14-16), zconv/grid_head/last = zero-initialised in the reference
(unet.py:26-30,142,226-228,529; poser_args.py:62-68; upscaler_02.py:49-51), siren* = SIREN uniform (siren.py:31-36).
"""
from typing import List, Tuple

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
