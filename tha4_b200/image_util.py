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


def poser_output_to_rgba_uint8_gpu(context, image: torch.Tensor, background=None, rint: bool = False) -> torch.Tensor:
    """The same conversion on the GPU (libtha4_b200: tha4_frame_to_srgb8): [B,4,H,W] or [4,H,W] device tensor ->
    [B,H,W,4] uint8 device tensor; a 512x512 frame then crosses PCIe as 1 MB instead of 4 MB of fp32.  `background`:
    None | 'green' | 'blue' | 'black' | 'white' (the puppeteers' output-background choices); rint=False truncates like
    the puppeteers' `.byte()`, rint=True rounds like convert_output_image_from_torch_to_numpy."""
    return context.frame_to_srgb8(image, background, rint)


def load_poser_image_gpu(context, path: str) -> torch.Tensor:
    """PNG -> [4,H,W] poser input with the colour conversion done on the GPU (tha4_rgba8_to_poser_image)."""
    import PIL.Image
    raw = torch.from_numpy(numpy.asarray(PIL.Image.open(path).convert('RGBA')).copy())
    return context.rgba8_to_poser_image(raw.to(context.device))
