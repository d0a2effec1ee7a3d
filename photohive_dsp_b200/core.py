"""Python front door with the reference's surface: ``get_report``, ``set_bounding_boxes``, ``Report``.

Mirrors core.py:23-119,219-228,388-515 of the reference (argument names, defaults, attribute names, the
JSON keys) on top of the CUDA build of libreport_data.so.  The matplotlib/tkinter viewers of the reference
(core.py:122-385) are out of scope; nothing GUI-related is imported here, so importing this module never
needs a display.  ``get_reports`` (plural) is the additive batch call.
"""
from __future__ import annotations

import ctypes
import json
import os
import time
from ctypes import POINTER
from types import SimpleNamespace

import numpy as np

from .batch import BatchReports, Context, PhotoHiveError, flat_layout, make_params
from .lib import lib
from .structures import Crop_Boundaries, Pixel_HSV
from .utils import array_to_image_rgb, hsv_to_rgb, image_pgm_to_pillow, pil_channels, pil_image_to_image_rgb

_VERBOSE = bool(os.environ.get("PHD_VERBOSE"))
_VIA_DOUBLES = bool(os.environ.get("PHD_GET_REPORT_VIA_DOUBLES"))  # force the reference's C entry point in get_report
_DEFAULT_DEVICE = int(os.environ.get("PHD_DEVICE", "0"))


class Report:
    """Python copy of one Full_Report_Data; owns the C object and frees it on deletion."""

    def __init__(self, report_ptr, height, width):
        data = report_ptr.contents
        self.data_ptr = report_ptr
        self.rgb_stats = data.rgb_stats.contents          # ctypes struct: Br Bg Bb Cr Cg Cb
        self.rgb_stats.height = height
        self.rgb_stats.width = width
        self.color_palette = self._convert_color_palette(data.color_palette)
        self.blur_profile = self._convert_blur_profile(data.blur_profile)
        self.blur_vectors = self._convert_blur_vectors(data.blur_vectors)
        self.average_saturation = data.average_saturation
        self.sharpnesses = self._convert_sharpnesses(data.sharpness)

    @staticmethod
    def _convert_sharpnesses(ptr):
        if not ptr:
            return []
        s = ptr.contents
        return [s.sharpness[i] for i in range(s.N)]

    @staticmethod
    def _convert_blur_vectors(ptr):
        group = ptr.contents
        return [SimpleNamespace(angle=group.blur_vectors[i].angle, magnitude=group.blur_vectors[i].magnitude)
                for i in range(group.len_vectors)]

    @staticmethod
    def _convert_color_palette(ptr):
        palette = ptr.contents
        entries = ctypes.cast(palette.averages, POINTER(Pixel_HSV * palette.N)).contents
        palette.colors = [hsv_to_rgb(e.h, e.s, e.v) for e in entries]
        palette.quantities = [palette.percentages[i] for i in range(palette.N)]
        return palette

    def _convert_blur_profile(self, ptr):
        c_profile = ptr.contents
        self.bp_ptr = c_profile
        rows = np.array(c_profile.get_bin_values(), dtype=np.float64)
        rows[np.isnan(rows)] = 0.0
        return SimpleNamespace(bins=rows.tolist())

    # -- images that need no GUI -----------------------------------------------------------------
    def generate_blur_profile_image(self):
        height, width = self.rgb_stats.height, self.rgb_stats.width
        pgm = lib.get_blur_profile_visual(ctypes.byref(self.bp_ptr), height, width)
        img = image_pgm_to_pillow(pgm, width, height)
        self.blur_profile_image = img.crop((0, 0, width // 2, height))
        return self.blur_profile_image

    def generate_color_palette_image(self, block=50):
        from PIL import Image, ImageDraw
        n = max(len(self.color_palette.colors), 1)
        per_row = int(np.ceil(np.sqrt(n)))
        img = Image.new("RGB", (per_row * block, ((n + per_row - 1) // per_row) * block), "black")
        draw = ImageDraw.Draw(img)
        for i, (color, share) in enumerate(zip(self.color_palette.colors, self.color_palette.quantities)):
            x, y = (i % per_row) * block, (i // per_row) * block
            draw.rectangle([x, y, x + block, y + block], fill=tuple(int(c) for c in color))
            draw.text((x + 4, y + block // 2 - 6), f"{share:.1%}", fill="black")
        self.color_palette_image = img
        return img

    def _no_viewer(self, *_a, **_k):
        raise NotImplementedError("the matplotlib/tkinter viewers of the reference are outside this package's scope; "
                                  "use generate_blur_profile_image()/generate_color_palette_image() and your own viewer")

    generate_blur_direction_frequency_response = _no_viewer
    display_color_palette_image = _no_viewer
    display_blur_profile = _no_viewer
    display_all = _no_viewer

    # -- export -----------------------------------------------------------------------------------
    def to_json(self):
        st = self.rgb_stats
        out = {"Height": st.height, "Width": st.width, "Average Saturation": self.average_saturation,
               "Red Brightness": st.Br, "Green Brightness": st.Bg, "Blue Brightness": st.Bb,
               "Red Contrast": st.Cr, "Green Contrast": st.Cg, "Blue Contrast": st.Cb}
        for i in range(10):
            out[f"Blur Vector {i+1} Angle"] = self.blur_vectors[i].angle
            out[f"Blur Vector {i+1} Magnitude"] = self.blur_vectors[i].magnitude
        colors, shares = self.color_palette.colors, self.color_palette.quantities
        for i in range(100):
            h, s, v = colors[i] if i < len(colors) else (0, 0, 0)
            out[f"Color {i+1} H"], out[f"Color {i+1} S"], out[f"Color {i+1} V"] = h, s, v
            out[f"Color {i+1} Percentage"] = shares[i] if i < len(colors) else 0
        for i in range(10):
            out[f"Sharpness {i+1}:"] = self.sharpnesses[i] if i < len(self.sharpnesses) else 0.0
        return json.dumps(out, indent=4)

    def __del__(self):
        ptr = getattr(self, "data_ptr", None)
        if ptr:
            lib.free_full_report(ctypes.byref(ptr))
            self.data_ptr = None


def get_report(pil_image, salient_characters=None,
               h_partitions=18, s_partitions=2, v_partitions=3,
               black_thresh=0.1, gray_thresh=0.1,
               coverage_thresh=0.95, linked_list_size=1000, downsample_rate=1,
               radius_partitions=40, angle_partitions=72,
               quantity_weight=0.1, saturation_value_weight=0.9,
               fft_streak_thresh=1.20, magnitude_thresh=0.3, blur_cutoff_ratio_denom=2):
    """One image through the drop-in C entry point; returns a Report, or None when the library refuses it.

    ``pil_image`` is a PIL image (as in the reference) or a uint8 [H,W,3] numpy array."""
    if salient_characters is None:
        boxes = POINTER(Crop_Boundaries)()
    elif isinstance(salient_characters, Crop_Boundaries):
        boxes = ctypes.byref(salient_characters)
    else:
        boxes = salient_characters
    params = (h_partitions, s_partitions, v_partitions, black_thresh, gray_thresh, coverage_thresh, linked_list_size,
              downsample_rate, radius_partitions, angle_partitions, quantity_weight, saturation_value_weight,
              fft_streak_thresh, magnitude_thresh, blur_cutoff_ratio_denom)
    t0 = time.time()
    ptr = None
    if not _VIA_DOUBLES and (salient_characters is None or isinstance(salient_characters, Crop_Boundaries)):
        ptr, height, width = _report_from_bytes(pil_image, salient_characters, params)
    if ptr is None:
        if isinstance(pil_image, np.ndarray):
            height, width = pil_image.shape[:2]
            image_rgb, _keep = array_to_image_rgb(pil_image)
        else:
            width, height = pil_image.width, pil_image.height
            image_rgb = pil_image_to_image_rgb(pil_image)
        ptr = lib.get_full_report_data(ctypes.byref(image_rgb), boxes, *params)
    if _VERBOSE:
        print(f"Elapsed time: {time.time() - t0} seconds")
    if not ptr:
        print("Failed to get report data")
        return None
    report = Report(ptr, height, width)
    report.magnitude_threshold = magnitude_thresh
    report.fft_streak_threshold = fft_streak_thresh
    report.blur_cutoff_ratio_denom = blur_cutoff_ratio_denom
    return report


def _report_from_bytes(image, crop, params):
    """8-bit shortcut of ``get_report``: the packed bytes go to the batch entry point (n = 1) and the flat record is
    turned into the same malloc'ed Full_Report_Data tree -- no 24 bytes per pixel of float64 planes on the way
    (utils.py:30-46 of the reference builds them only because its C side wants doubles).  Results are identical to the
    C entry point's (tests/test_gpu_parity.py).  Returns (None, h, w) whenever the reference's own entry point should
    decide instead: refused sizes, boxes outside the image, an empty box list."""
    if isinstance(image, np.ndarray):
        arr = image
    else:
        arr = pil_channels(image)  # first three channels of np.array(image), like the reference (utils.py:30-33)
    if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
        return None, 0, 0          # 16-bit / float images: the planes of doubles go to the C entry point
    arr = np.ascontiguousarray(arr)
    height, width = arr.shape[:2]
    nb, barr = 0, None
    if crop is not None:
        nb = int(crop.N)
        if nb <= 0:
            return None, height, width
        barr = np.array([[crop.top[i], crop.bottom[i], crop.left[i], crop.right[i]] for i in range(nb)], np.int32)
        if (barr < 0).any() or (barr[:, :2] > height).any() or (barr[:, 2:] > width).any():
            return None, height, width
    ctx = _contexts.get(_DEFAULT_DEVICE)
    if ctx is None:
        ctx = _contexts[_DEFAULT_DEVICE] = Context(_DEFAULT_DEVICE)
    names = ("h_partitions", "s_partitions", "v_partitions", "black_thresh", "gray_thresh", "coverage_thresh",
             "linked_list_size", "downsample_rate", "radius_partitions", "angle_partitions", "quantity_weight",
             "saturation_value_weight", "fft_streak_thresh", "magnitude_thresh", "blur_cutoff_ratio_denom")
    p = make_params(**dict(zip(names, params)))
    try:
        lay = flat_layout(p, nb)
        rec = np.empty((1, lay.record_bytes), np.uint8)
        ctx.get_reports_raw(arr.ctypes.data, 1, width, height, width * height * 3, p, rec.ctypes.data,
                            boxes_ptr=None if barr is None else barr.ctypes.data, max_boxes=nb)
    except PhotoHiveError as e:
        if e.code in (1, 2):        # refused by the reference's pre-checks / bad parameters:
            return None, height, width  # let the C entry point print the reference's message
        raise                       # a CUDA fault or an unsupported size is not something to retry
    ptr = lib.phd_flat_to_full_report(rec.ctypes.data, ctypes.byref(lay))
    return (ptr if ptr else None), height, width


def set_bounding_boxes(bounding_boxes):
    """list of dicts with 'top', 'bottom', 'left', 'right' -> Crop_Boundaries (core.py:489-515)."""
    n = len(bounding_boxes)
    cols = {k: (ctypes.c_int * n)(*[int(b[k]) for b in bounding_boxes]) for k in ("top", "bottom", "left", "right")}
    cb = Crop_Boundaries(N=n, top=cols["top"], bottom=cols["bottom"], left=cols["left"], right=cols["right"])
    cb._keepalive = cols
    return cb


_contexts: dict[int, Context] = {}


def get_reports(images, boxes=None, device: int = 0, **params) -> BatchReports:
    """Additive batch call: uint8 [n,H,W,3] (numpy host array or torch CUDA tensor) -> BatchReports."""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx.get_reports(images, boxes=boxes, params=make_params(**params))
