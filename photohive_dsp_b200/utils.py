"""Image container conversions for the drop-in API (counterpart of the reference's utils.py)."""
import ctypes as C

import numpy as np

from .structures import Image_RGB

_dp = C.POINTER(C.c_double)


def hsv_to_rgb(h, s, v):
    """Palette entry (h in degrees, s, v in [0,1]) -> integer (r, g, b); same sector convention as utils.py:8-28."""
    chroma = v * s
    second = chroma * (1 - abs((h / 60) % 2 - 1))
    base = v - chroma
    sector = next(i for i, edge in enumerate((60, 120, 180, 240, 300, float("inf"))) if h < edge or i == 5)
    r, g, b = [(chroma, second, 0), (second, chroma, 0), (0, chroma, second),
               (0, second, chroma), (second, 0, chroma), (chroma, 0, second)][sector]
    return int((r + base) * 255), int((g + base) * 255), int((b + base) * 255)


def pil_channels(pil_image) -> np.ndarray:
    """The array the reference works on (utils.py:30-33): ``np.array(pil_image)``, first three channels.  No mode
    conversion: RGBA drops alpha, CMYK / YCbCr / LAB pass their first three channels through as the reference does, and a
    single-channel image ('L', 'P', 'I;16' ...) raises the same IndexError (the array is 2-D)."""
    arr = np.array(pil_image)
    arr[:, :, 2]  # IndexError for 2-D (single channel) and two-channel arrays, exactly where the reference raises it
    return arr[:, :, 0:3]


def pil_image_to_image_rgb(pil_image):
    """PIL image -> Image_RGB of three contiguous float64 planes holding value/255.0 (utils.py:30-46).
    The planes are parked on the PIL object so they outlive the C call."""
    width, height = pil_image.size
    arr = pil_channels(pil_image)
    planes = [np.ascontiguousarray(arr[:, :, c] / 255.0, dtype=np.float64).ravel() for c in range(3)]
    ptrs = [p.ctypes.data_as(_dp) for p in planes]
    pil_image._phd_planes = planes
    pil_image.r_ctypes, pil_image.g_ctypes, pil_image.b_ctypes = ptrs
    return Image_RGB(height=height, width=width, r=ptrs[0], g=ptrs[1], b=ptrs[2])


def array_to_image_rgb(rgb_u8: np.ndarray):
    """uint8 [H,W,3] -> (Image_RGB, keepalive)."""
    h, w, _ = rgb_u8.shape
    planes = [np.ascontiguousarray(rgb_u8[:, :, c].astype(np.float64) / 255.0).ravel() for c in range(3)]
    ptrs = [p.ctypes.data_as(_dp) for p in planes]
    return Image_RGB(height=h, width=w, r=ptrs[0], g=ptrs[1], b=ptrs[2]), planes


def _plane(ptr, width, height):
    return np.ctypeslib.as_array(ptr, shape=(height * width,)).reshape(height, width)


def image_rgb_to_pillow(image_rgb_ptr, width, height):
    from PIL import Image
    img = image_rgb_ptr.contents
    stack = np.stack([_plane(img.r, width, height), _plane(img.g, width, height), _plane(img.b, width, height)], -1)
    return Image.fromarray(np.clip(stack * 255, 0, 255).astype(np.uint8), "RGB")


def image_pgm_to_pillow(image_pgm_ptr, width, height):
    from PIL import Image
    data = _plane(image_pgm_ptr.contents.data, width, height)
    return Image.fromarray(np.clip(data * 255, 0, 255).astype(np.uint8), "L")
