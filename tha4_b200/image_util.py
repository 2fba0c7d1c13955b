"""Image I/O on either side of the poser: PNG -> poser input tensor and poser output -> displayable RGBA.

Restates src/tha4/shion/base/image_util.py of the reference: sRGB -> linear (:10-12), premultiplied alpha
(:147-148), [0,1] -> [-1,1] (:149), HWC -> CHW (:152-162); and the inverse used by the GUIs
(src/tha4/image_util.py:56-58, shion/base/image_util.py:90-108)."""
import numpy
import torch


def numpy_srgb_to_linear(x):
    x = numpy.clip(x, 0.0, 1.0)
    return numpy.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


def numpy_linear_to_srgb(x):
    x = numpy.clip(x, 0.0, 1.0)
    return numpy.where(x <= 0.003130804953560372, x * 12.92, 1.055 * (x ** (1.0 / 2.4)) - 0.055)


def load_poser_image(path: str) -> torch.Tensor:
    """[4,H,W] float32 in [-1,1], linear RGB premultiplied by alpha (extract_pytorch_image_from_filelike)."""
    import PIL.Image
    pil = PIL.Image.open(path).convert('RGBA')
    raw = numpy.asarray(pil, dtype=numpy.float32) / 255.0
    raw[:, :, 0:3] = numpy_srgb_to_linear(raw[:, :, 0:3])
    raw[:, :, 0:3] = raw[:, :, 0:3] * raw[:, :, 3:4]
    return torch.from_numpy(numpy.ascontiguousarray((raw * 2.0 - 1.0).transpose(2, 0, 1))).float()


def poser_output_to_rgba_uint8(image: torch.Tensor) -> numpy.ndarray:
    """[4,H,W] poser output -> HxWx4 uint8 sRGB (pytorch_rgba_to_numpy_image + convert_output_image_from_torch_to_numpy)."""
    x = (image.detach().float().cpu().numpy().transpose(1, 2, 0) + 1.0) / 2.0
    rgb = numpy_linear_to_srgb(x[:, :, 0:3])
    a = numpy.clip(x[:, :, 3:4], 0.0, 1.0)
    return numpy.uint8(numpy.rint(numpy.concatenate([rgb, a], axis=2) * 255.0))
