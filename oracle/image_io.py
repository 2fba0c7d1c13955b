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


def linear_to_srgb_torch(x: torch.Tensor) -> torch.Tensor:
    """shion/base/image_util.py:30-32 (torch_linear_to_srgb)."""
    x = torch.clip(x, 0.0, 1.0)
    return torch.where(torch.le(x, 0.003130804953560372), x * 12.92, 1.055 * (x ** (1.0 / 2.4)) - 0.055)


BACKGROUND_COLOURS = {0: None, 1: (0.0, 1.0, 0.0), 2: (0.0, 0.0, 1.0), 3: (0.0, 0.0, 0.0), 4: (1.0, 1.0, 1.0)}


def frame_to_srgb8(frame: torch.Tensor, background: int = 0, rint: bool = False) -> torch.Tensor:
    """The display conversion of the puppeteer apps, restated op by op from
    app/character_model_ifacialmocap_puppeteer.py:325-349 (+ blend_with_background :377-381 and convert_linear_to_srgb,
    tha4/image_util.py:61-63): [4,H,W] poser output in [-1,1] -> [H,W,4] uint8.  rint=False is the puppeteer's
    `.byte()` (truncation); rint=True is convert_output_image_from_torch_to_numpy's numpy.rint (tha4/image_util.py:56)."""
    out = torch.clip((frame.float() + 1.0) / 2.0, 0.0, 1.0)
    out = torch.cat([linear_to_srgb_torch(out[0:3]), out[3:4]], dim=0)
    colour = BACKGROUND_COLOURS[background]
    if colour is not None:
        bg = torch.zeros(4, out.shape[1], out.shape[2])
        bg[3] = 1.0
        for c in range(3):
            bg[c] = colour[c]
        alpha, col = out[3:4], out[0:3]
        out = torch.cat([col * alpha + (1.0 - alpha) * bg[0:3], bg[3:4]], dim=0)
    c, h, w = out.shape
    out = 255.0 * torch.transpose(out.reshape(c, h * w), 0, 1).reshape(h, w, c)
    return torch.from_numpy(numpy.uint8(numpy.rint(out.numpy()))) if rint else out.byte()
