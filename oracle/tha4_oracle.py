"""CPU oracle for the THA4 poser hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A from-scratch, functional restatement (plain PyTorch fp32 on the CPU) of the
reference's per-frame forward pass: the five teacher networks, the two SIREN
students and the three poser "modes" that chain them.  Every function cites the
reference file:line it restates (paths relative to /root/reference/src/tha4).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (tha4_b200) never does.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md F2), so this restatement is pinned against *outputs of the reference
itself*: oracle/make_golden.py imports the real reference from /root/reference
in the build container, runs it on seeded weights/inputs, and commits
sub-sampled outputs under tests/golden/; tests/test_oracle_pinned.py checks this
file against them (and, when /root/reference is present, against the live
reference on full tensors).

All functions take `sd`, a state_dict in the reference's on-disk key layout
(shion/core/load_save.py:6-14), and a key `prefix`.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]

NUM_EYEBROW_PARAMS = 12   # poser/modes/mode_07.py:42
NUM_FACE_PARAMS = 27      # poser/modes/mode_07.py:43
NUM_ROTATION_PARAMS = 6   # poser/modes/mode_07.py:44


# --------------------------------------------------------------------------------------
# shared image ops
# --------------------------------------------------------------------------------------
def base_grid(n: int, h: int, w: int, dtype=torch.float32) -> Tensor:
    """affine_grid(identity, align_corners=False) -> [n,h,w,2], channel 0 = x.
    nn/image_processing_util.py:17-22."""
    theta = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=dtype).unsqueeze(0).repeat(n, 1, 1)
    return F.affine_grid(theta, [n, 1, h, w], align_corners=False)


def apply_grid_change(grid_change: Tensor, image: Tensor) -> Tensor:
    """nn/image_processing_util.py:13-24 and :33-54 (GridChangeApplier.apply):
    [n,2,h,w] offsets -> [n,h,w,2]; grid = base + offsets; bilinear, border, align_corners=False."""
    n, c, h, w = image.shape
    gc = grid_change.reshape(n, 2, h * w).transpose(1, 2).reshape(n, h, w, 2)
    grid = base_grid(n, h, w, grid_change.dtype) + gc
    return F.grid_sample(image, grid, mode='bilinear', padding_mode='border', align_corners=False)


def apply_color_change(alpha: Tensor, color_change: Tensor, image: Tensor) -> Tensor:
    """nn/image_processing_util.py:57-58."""
    return color_change * alpha + image * (1 - alpha)


def apply_rgb_change(alpha: Tensor, color_change: Tensor, image: Tensor) -> Tensor:
    """nn/image_processing_util.py:6-10 -- blends RGB only, keeps the image's own alpha."""
    rgb = color_change[:, 0:3] * alpha + image[:, 0:3] * (1 - alpha)
    return torch.cat([rgb, image[:, 3:4]], dim=1)


# --------------------------------------------------------------------------------------
# encoder-decoder trunk (EyebrowDecomposer00 / EyebrowMorphingCombiner00 / FaceMorpher08)
# --------------------------------------------------------------------------------------
def _inorm(sd: SD, key: str, x: Tensor) -> Tensor:
    """InstanceNorm2d(affine=True), eps 1e-5, biased variance, no running stats. nn/normalization.py:94-95."""
    return F.instance_norm(x, weight=sd[key + '.weight'], bias=sd[key + '.bias'], eps=1e-5)


def _conv_in_relu(sd: SD, p: str, x: Tensor, stride: int, pad: int) -> Tensor:
    """Sequential(conv(no bias), InstanceNorm2d, ReLU). nn/conv.py:103-113 / :127-147."""
    return F.relu(_inorm(sd, p + '.1', F.conv2d(x, sd[p + '.0.weight'], None, stride, pad)))


def _deconv_in_relu(sd: SD, p: str, x: Tensor) -> Tensor:
    """Sequential(ConvTranspose2d 4x4 s2 p1 (no bias), InstanceNorm2d, ReLU). nn/conv.py:164-177."""
    return F.relu(_inorm(sd, p + '.1', F.conv_transpose2d(x, sd[p + '.0.weight'], None, 2, 1)))


def _resnet_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """x + [conv3, IN, ReLU, conv3, IN](x). nn/resnet_block.py:52-61,64-67."""
    h = F.conv2d(x, sd[p + '.resnet_path.0.weight'], None, 1, 1)
    h = F.relu(_inorm(sd, p + '.resnet_path.1', h))
    h = F.conv2d(h, sd[p + '.resnet_path.3.weight'], None, 1, 1)
    h = _inorm(sd, p + '.resnet_path.4', h)
    return x + h


def encoder_decoder_trunk(sd: SD, p: str, image: Tensor, pose: Optional[Tensor],
                          num_levels: int = 4, num_bottleneck: int = 6) -> Tensor:
    """nn/common/poser_encoder_decoder_00.py:99-121 and nn/face_morpher/face_morpher_08.py:158-168.
    Returns the last (full-resolution, 64-channel) feature map."""
    f = _conv_in_relu(sd, p + 'downsample_blocks.0', image, 1, 1)
    for i in range(1, num_levels):
        f = _conv_in_relu(sd, p + 'downsample_blocks.%d' % i, f, 2, 1)
    if pose is not None:
        n, c = pose.shape
        s = f.shape[2]
        f = torch.cat([f, pose.view(n, c, 1, 1).repeat(1, 1, s, s)], dim=1)
    f = _conv_in_relu(sd, p + 'bottleneck_blocks.0', f, 1, 1)
    for i in range(1, num_bottleneck):
        f = _resnet_block(sd, p + 'bottleneck_blocks.%d' % i, f)
    for i in range(num_levels - 1):
        f = _deconv_in_relu(sd, p + 'upsample_blocks.%d' % i, f)
    return f


def _head(sd: SD, key: str, feature: Tensor) -> Tensor:
    return F.conv2d(feature, sd[key + '.weight'], sd.get(key + '.bias'), 1, 1)


def eyebrow_decomposer(sd: SD, image: Tensor) -> List[Tensor]:
    """nn/eyebrow_decomposer/eyebrow_decomposer_00.py:46-64."""
    feature = encoder_decoder_trunk(sd, 'body.', image, None)
    bg_alpha = torch.sigmoid(_head(sd, 'background_layer_alpha.0', feature))
    bg_color = torch.tanh(_head(sd, 'background_layer_color_change.0', feature))
    bg_layer = apply_color_change(bg_alpha, bg_color, image)
    eb_alpha = torch.sigmoid(_head(sd, 'eyebrow_layer_alpha.0', feature))
    eb_color = torch.tanh(_head(sd, 'eyebrow_layer_color_change.0', feature))
    eb_layer = apply_color_change(eb_alpha, image, eb_color)   # note the swapped roles (:57)
    return [eb_layer, eb_alpha, eb_color, bg_layer, bg_alpha, bg_color]


def eyebrow_morphing_combiner(sd: SD, background_layer: Tensor, eyebrow_layer: Tensor, pose: Tensor) -> List[Tensor]:
    """nn/eyebrow_morphing_combiner/eyebrow_morphing_combiner_00.py:47-72."""
    feature = encoder_decoder_trunk(sd, 'body.', torch.cat([background_layer, eyebrow_layer], dim=1), pose)
    grid_change = _head(sd, 'morphed_eyebrow_layer_grid_change', feature)
    alpha = torch.sigmoid(_head(sd, 'morphed_eyebrow_layer_alpha.0', feature))
    color = torch.tanh(_head(sd, 'morphed_eyebrow_layer_color_change.0', feature))
    warped = apply_grid_change(grid_change, eyebrow_layer)
    morphed = apply_color_change(alpha, color, warped)
    combine_alpha = torch.sigmoid(_head(sd, 'combine_alpha.0', feature))
    eyebrow_image = apply_rgb_change(combine_alpha, morphed, background_layer)
    eyebrow_image_no_combine_alpha = apply_rgb_change((morphed[:, 3:4] + 1.0) / 2.0, morphed, background_layer)
    return [eyebrow_image, combine_alpha, eyebrow_image_no_combine_alpha, morphed, alpha, color, warped, grid_change]


def face_morpher(sd: SD, image: Tensor, pose: Tensor) -> List[Tensor]:
    """nn/face_morpher/face_morpher_08.py:158-193 (output_iris_mouth_grid_change=True, mode_07.py:205)."""
    feature = encoder_decoder_trunk(sd, '', image, pose)
    im_grid_change = _head(sd, 'iris_mouth_grid_change', feature)
    im_image_0 = apply_grid_change(im_grid_change, image)
    im_color = torch.tanh(_head(sd, 'iris_mouth_color_change.0', feature))
    im_alpha = torch.sigmoid(_head(sd, 'iris_mouth_alpha.0', feature))
    im_image_1 = apply_color_change(im_alpha, im_color, im_image_0)
    eye_color = torch.tanh(_head(sd, 'eye_color_change.0', feature))
    eye_alpha = torch.sigmoid(_head(sd, 'eye_alpha.0', feature))
    output_image = apply_color_change(eye_alpha, eye_color, im_image_1)
    return [output_image, eye_alpha, eye_color, im_image_1, im_alpha, im_color, im_image_0, im_grid_change]


# --------------------------------------------------------------------------------------
# diffusion-style U-Net (Morpher00 / Upscaler02)
# --------------------------------------------------------------------------------------
def _gn32(sd: SD, key: str, x: Tensor) -> Tensor:
    """GroupNorm(min(32,C), C), eps 1e-5. nn/common/unet.py:65-66."""
    c = x.shape[1]
    return F.group_norm(x, min(32, c), sd[key + '.weight'], sd[key + '.bias'], eps=1e-5)


def _scaleshift(x: Tensor, ss: Tensor) -> Tensor:
    """x * (1 + scale) + shift, scale/shift = chunk(ss, 2, dim=1). nn/common/unet.py:90-97."""
    scale, shift = torch.chunk(ss.reshape(ss.shape[0], ss.shape[1], 1, 1), 2, dim=1)
    return x * (1.0 + scale) + shift


def _res_block(sd: SD, p: str, x: Tensor, cond0: Tensor, cond1: Tensor, mode: str = 'same') -> Tensor:
    """nn/common/unet.py:154-165.  mode: 'same' | 'up' (nearest x2, :46) | 'down' (AvgPool2d(2), :58)."""
    def resample(t):
        if mode == 'up':
            return F.interpolate(t, scale_factor=2, mode='nearest')
        if mode == 'down':
            return F.avg_pool2d(t, 2, 2)
        return t
    h = F.conv2d(resample(F.silu(_gn32(sd, p + '.norm0', x))), sd[p + '.conv0.weight'], sd[p + '.conv0.bias'], 1, 1)
    h = _gn32(sd, p + '.norm1', h)
    h = _scaleshift(h, F.linear(F.silu(cond0), sd[p + '.cond0_layers.1.weight'], sd[p + '.cond0_layers.1.bias']))
    h = _scaleshift(h, F.linear(F.silu(cond1), sd[p + '.cond1_layers.1.weight'], sd[p + '.cond1_layers.1.bias']))
    h = F.conv2d(F.silu(h), sd[p + '.conv1.weight'], sd[p + '.conv1.bias'], 1, 1)
    xr = resample(x)
    if (p + '.skip.weight') in sd:
        xr = F.conv2d(xr, sd[p + '.skip.weight'], sd[p + '.skip.bias'])
    return xr + h


def _attention_block(sd: SD, p: str, x: Tensor, num_heads: int = 8) -> Tensor:
    """nn/common/unet.py:230-239 with qkv_attention ('new order', :192-202)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(_gn32(sd, p + '.norm', x), sd[p + '.qkv.weight'], sd[p + '.qkv.bias']).reshape(b, 3 * c, hh * ww)
    L = hh * ww
    ch = c // num_heads
    q, k, v = qkv.chunk(3, dim=1)
    scale = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.einsum('bct,bcs->bts', (q * scale).reshape(b * num_heads, ch, L), (k * scale).reshape(b * num_heads, ch, L))
    w = torch.softmax(w, dim=-1)
    a = torch.einsum('bts,bcs->bct', w, v.reshape(b * num_heads, ch, L)).reshape(b, c, hh, ww)
    return x + F.conv2d(a, sd[p + '.conv.weight'], sd[p + '.conv.bias'])


def _timestep_embedding_zero(n: int, channels: int) -> Tensor:
    """nn/common/unet.py:365-376 evaluated at t = 0 (morpher_00.py:51, upscaler_02.py:78):
    cat(cos(0), sin(0)) = [1]*half ++ [0]*half."""
    half = channels // 2
    return torch.cat([torch.ones(n, half), torch.zeros(n, half)], dim=1)


def unet(sd: SD, p: str, x: Tensor, cond: Tensor, model_channels: int, mults: List[int],
         use_attention: List[bool], first_conv_addition: Optional[Tensor] = None) -> Tensor:
    """nn/common/unet.py:531-546 (Unet.forward) and :642-658 (UnetWithFirstConvAddition.forward);
    structure per :454-529 with num_res_blocks_per_level=1, num_middle_res_blocks=4, resample_with_res_block."""
    n = x.shape[0]
    t_emb = _timestep_embedding_zero(n, model_channels)
    t_emb = F.linear(t_emb, sd[p + 'time_embed.1.weight'], sd[p + 'time_embed.1.bias'])
    t_emb = F.linear(F.silu(t_emb), sd[p + 'time_embed.3.weight'], sd[p + 'time_embed.3.bias'])
    c_emb = F.linear(cond, sd[p + 'cond_embed.0.weight'], sd[p + 'cond_embed.0.bias'])
    c_emb = F.linear(F.silu(c_emb), sd[p + 'cond_embed.2.weight'], sd[p + 'cond_embed.2.bias'])
    h = F.conv2d(x, sd[p + 'first_conv.weight'], sd[p + 'first_conv.bias'], 1, 1)
    if first_conv_addition is not None:
        h = h + first_conv_addition
    hs = [h]
    num_levels = len(mults)
    for i in range(num_levels):
        bp = p + 'down_blocks.%d' % i
        h = _res_block(sd, bp + '.res_blocks.0', hs[-1], t_emb, c_emb)
        if use_attention[i]:
            h = _attention_block(sd, bp + '.attention_blocks.0', h)
        hs.append(h)
        if i < num_levels - 1:
            hs.append(_res_block(sd, bp + '.downsample', h, t_emb, c_emb, 'down'))
    h = hs[-1]
    for j in range(7):  # Res, Attn, Res, Attn, Res, Attn, Res (:481-498)
        mp = p + 'middle_blocks.%d' % j
        h = _res_block(sd, mp, h, t_emb, c_emb) if j % 2 == 0 else _attention_block(sd, mp + '.module', h)
    for bi, i in enumerate(reversed(range(num_levels))):
        bp = p + 'up_blocks.%d' % bi
        for r in range(2):
            h = _res_block(sd, bp + '.resnet_blocks.%d' % r, torch.cat([h, hs.pop()], dim=1), t_emb, c_emb)
            if use_attention[i]:
                h = _attention_block(sd, bp + '.attention_blocks.%d' % r, h)
        if i > 0:
            h = _res_block(sd, bp + '.upsample', h, t_emb, c_emb, 'up')
    assert len(hs) == 0
    h = F.silu(_gn32(sd, p + 'last.0', h))
    return F.conv2d(h, sd[p + 'last.2.weight'], sd[p + 'last.2.bias'], 1, 1)


MORPHER_UNET = dict(model_channels=64, mults=[1, 2, 4, 4, 4], use_attention=[False] * 4 + [True])    # mode_07.py:210-226
UPSCALER_UNET = dict(model_channels=32, mults=[1, 2, 4, 8, 8, 8], use_attention=[False] * 5 + [True])  # mode_07.py:241-257


def _unet_tail(body_output: Tensor, image: Tensor) -> List[Tensor]:
    """morpher_00.py:53-66 / upscaler_02.py:82-96: split 7 channels, sigmoid alpha, warp, blend."""
    direct = body_output[:, 0:4]
    grid_change = body_output[:, 4:6]
    alpha = torch.sigmoid(body_output[:, 6:7])
    warped = apply_grid_change(grid_change, image)
    merged = apply_color_change(alpha, direct, warped)
    return [merged, alpha, warped, grid_change, direct]


def morpher_00(sd: SD, image: Tensor, pose: Tensor) -> List[Tensor]:
    """nn/morpher/morpher_00.py:42-66."""
    return _unet_tail(unet(sd, 'body.', image, pose, **MORPHER_UNET), image)


def upscaler_02(sd: SD, rest_image: Tensor, coarse_posed_image: Tensor, coarse_grid_change: Tensor,
                pose: Tensor) -> List[Tensor]:
    """nn/upscaler/upscaler_02.py:59-96."""
    warped_image = apply_grid_change(coarse_grid_change, rest_image)
    feature = torch.cat([coarse_posed_image, warped_image, coarse_grid_change], dim=1)
    addition = F.conv2d(feature, sd['coarse_image_conv.weight'], sd['coarse_image_conv.bias'], 1, 1)
    body_output = unet(sd, 'body.', rest_image, pose, first_conv_addition=addition, **UPSCALER_UNET)
    return _unet_tail(body_output, rest_image)


# --------------------------------------------------------------------------------------
# SIREN students
# --------------------------------------------------------------------------------------
def _position_grid(n: int, size: int) -> Tensor:
    """siren_morpher_03.py:92-99 / siren_face_morpher_00.py:38-44: [n,2,size,size], channel 0 = x."""
    return base_grid(1, size, size).reshape(1, size * size, 2).transpose(1, 2).reshape(1, 2, size, size).repeat(n, 1, 1, 1)


def _sine_layer(sd: SD, p: str, x: Tensor) -> Tensor:
    """sin(30 * conv1x1(x)). nn/siren/vanilla/siren.py:38-39."""
    return torch.sin(30.0 * F.conv2d(x, sd[p + '.linear.weight'], sd[p + '.linear.bias']))


def siren_face_morpher(sd: SD, pose: Tensor, image_size: int = 128, num_sine_layers: int = 8) -> Tensor:
    """nn/siren/face_morpher/siren_face_morpher_00.py:34-51 + vanilla/siren.py:84-91 (no output nonlinearity)."""
    n, p = pose.shape
    x = torch.cat([_position_grid(n, image_size), pose.view(n, p, 1, 1).repeat(1, 1, image_size, image_size)], dim=1)
    for i in range(num_sine_layers):
        x = _sine_layer(sd, 'siren.sine_layers.%d' % i, x)
    return F.conv2d(x, sd['siren.last_linear.weight'], sd['siren.last_linear.bias'])


def siren_morpher_03(sd: SD, image: Tensor, pose: Tensor, level_sizes=(128, 256, 512)) -> List[Tensor]:
    """nn/siren/morpher/siren_morpher_03.py:107-139."""
    n, p = pose.shape
    x = None
    for i, size in enumerate(level_sizes):
        pp = torch.cat([_position_grid(n, size), pose.view(n, p, 1, 1).repeat(1, 1, size, size)], dim=1)
        if i == 0:
            x = pp
        else:
            x = torch.cat([F.interpolate(x, size=(size, size), mode='bilinear'), pp], dim=1)
        for j in range(3):
            x = _sine_layer(sd, 'siren_layers.%d.%d' % (i, j), x)
    out = F.conv2d(x, sd['last_linear.weight'], sd['last_linear.bias'])
    grid_change = out[:, 0:2]
    alpha = out[:, 2:3]            # raw, no sigmoid (:128)
    color_change = out[:, 3:]
    warped = apply_grid_change(grid_change, image)
    blended = (1 - alpha) * warped + alpha * color_change
    return [blended, alpha, color_change, warped, grid_change]


# --------------------------------------------------------------------------------------
# poser modes
# --------------------------------------------------------------------------------------
def _promote(image: Tensor, pose: Tensor):
    """poser/general_poser_02.py:66-69."""
    if image.dim() == 3:
        image = image.unsqueeze(0)
    if pose.dim() == 1:
        pose = pose.unsqueeze(0)
    return image, pose


def mode_12_outputs(sds: Dict[str, SD], image: Tensor, pose: Tensor, eyebrow_morphed_image_index: int = 2,
                    cached_decomposer_output: Optional[List[Tensor]] = None) -> List[Tensor]:
    """poser/modes/mode_12.py:66-94: face morpher(8) + combiner(8) + decomposer(6) = 22 tensors."""
    image, pose = _promote(image, pose)
    dec = cached_decomposer_output
    if dec is None:
        dec = eyebrow_decomposer(sds['eyebrow_decomposer'], image[:, :, 64:192, 192:320])
    comb = eyebrow_morphing_combiner(sds['eyebrow_morphing_combiner'], dec[3], dec[0], pose[:, :NUM_EYEBROW_PARAMS])
    face_in = image[:, :, 32:224, 160:352].clone()
    face_in[:, :, 32:160, 32:160] = comb[eyebrow_morphed_image_index]
    face = face_morpher(sds['face_morpher'], face_in, pose[:, NUM_EYEBROW_PARAMS:NUM_EYEBROW_PARAMS + NUM_FACE_PARAMS])
    return face + comb + dec


def mode_07_outputs(sds: Dict[str, SD], image: Tensor, pose: Tensor, eyebrow_morphed_image_index: int = 2,
                    cached_decomposer_output: Optional[List[Tensor]] = None) -> List[Tensor]:
    """poser/modes/mode_07.py:72-132: upscaler(5) + face_morphed_full(1) + body(5) + face(8) + combiner(8)
    + decomposer(6) = 33 tensors."""
    image, pose = _promote(image, pose)
    fcd = mode_12_outputs(sds, image, pose, eyebrow_morphed_image_index, cached_decomposer_output)
    face, comb, dec = fcd[0:8], fcd[8:16], fcd[16:22]
    full = image.clone()
    full[:, :, 32:224, 160:352] = face[0]
    half = F.interpolate(full, size=(256, 256), mode='bilinear', align_corners=False)
    rot = pose[:, NUM_EYEBROW_PARAMS + NUM_FACE_PARAMS:]
    body = morpher_00(sds['body_morpher'], half, rot)
    coarse_posed = F.interpolate(body[0], size=(512, 512), mode='bilinear')
    coarse_grid = F.interpolate(body[3], size=(512, 512), mode='bilinear')
    up = upscaler_02(sds['upscaler'], full, coarse_posed, coarse_grid, rot)
    return up + [full] + body + face + comb + dec


def mode_14_outputs(sds: Dict[str, SD], image: Tensor, pose: Tensor) -> List[Tensor]:
    """poser/modes/mode_14.py:52-90: body(5) + [face] = 6 tensors."""
    image, pose = _promote(image, pose)
    face = siren_face_morpher(sds['face_morpher'], pose[:, 0:39])
    body_in = image.clone()
    body_in[:, :, 80:208, 192:320] = face
    body = siren_morpher_03(sds['body_morpher'], body_in, pose)
    return body + [face]
