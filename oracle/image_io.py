"""PNG -> poser input tensor, restated for the oracle / fixtures  --  TEST INFRASTRUCTURE.

Follows shion/base/image_util.py:127-149 (sRGB->linear :10-12, premultiply alpha :147-148, x*2-1 :149)
and the HWC->CHW step of :152-162."""
import numpy
import torch


def srgb_to_linear(x):
    x = numpy.clip(x, 0.0, 1.0)
    return numpy.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


def load_rgba_png(path: str) -> torch.Tensor:
    import PIL.Image
    pil = PIL.Image.open(path).convert('RGBA')
    raw = numpy.asarray(pil, dtype=numpy.float32) / 255.0
    raw[:, :, 0:3] = srgb_to_linear(raw[:, :, 0:3])
    raw[:, :, 0:3] = raw[:, :, 0:3] * raw[:, :, 3:4]
    img = raw * 2.0 - 1.0
    return torch.from_numpy(numpy.ascontiguousarray(img.transpose(2, 0, 1))).float()
