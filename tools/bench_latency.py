"""Single-image latency of the drop-in call and of the batch entry point with n = 1 (BASELINE configs 1 and 2).

  python tools/bench_latency.py           # on a GPU box

Prints, per shape: get_full_report_data (the reference's C entry point: three planes of doubles in pageable host
memory -> malloc'ed report tree), phd_get_reports_u8 with one packed 8-bit image in pinned host memory, and the same
with the image already resident on the device.  Wall clock around the blocking calls, median of the repeats.
"""
import ctypes
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from photohive_dsp_b200.core import set_bounding_boxes  # noqa: E402
from photohive_dsp_b200.lib import lib  # noqa: E402
from photohive_dsp_b200.structures import Crop_Boundaries  # noqa: E402
from photohive_dsp_b200.utils import array_to_image_rgb  # noqa: E402
from tools.synth import Generator  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)
DEFAULTS = (18, 2, 3, 0.1, 0.1, 0.95, 1000, 1, 40, 72, 0.1, 0.9, 1.20, 0.3, 2)


def med(f, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts)


def run(name, W, H, boxes, reps=15):
    img_dev = Generator(W, H, dev).batch(1)
    img = img_dev.cpu().numpy()[0]
    image_rgb, keep = array_to_image_rgb(img)
    cb = set_bounding_boxes(boxes) if boxes else None
    bptr = ctypes.byref(cb) if cb else ctypes.POINTER(Crop_Boundaries)()

    def dropin():
        p = lib.get_full_report_data(ctypes.byref(image_rgb), bptr, *DEFAULTS)
        assert p
        lib.free_full_report(ctypes.byref(p))

    params = make_params()
    nb = len(boxes) if boxes else 0
    lay = flat_layout(params, nb)
    barr = np.ascontiguousarray([[[b["top"], b["bottom"], b["left"], b["right"]] for b in boxes]], np.int32) if boxes else None
    bkw = {} if barr is None else dict(boxes_ptr=barr.ctypes.data, max_boxes=nb)
    pinned = torch.from_numpy(img[None]).pin_memory()
    rec_h = torch.empty((1, lay.record_bytes), dtype=torch.uint8).pin_memory()
    rec_d = torch.empty((1, lay.record_bytes), dtype=torch.uint8, device=dev)

    def batch_host():
        ctx.get_reports_raw(pinned.data_ptr(), 1, W, H, W * H * 3, params, rec_h.data_ptr(), **bkw)

    def batch_dev():
        ctx.get_reports_raw(img_dev.data_ptr(), 1, W, H, W * H * 3, params, rec_d.data_ptr(), **bkw)
        torch.cuda.synchronize()

    for f in (dropin, batch_host, batch_dev):
        for _ in range(3):
            f()
    a, b, c = med(dropin, reps), med(batch_host, reps), med(batch_dev, reps)
    ms, _ = ctx.last_timing()
    print(f"{name}: get_full_report_data {a[0]:.2f} ms (min {a[1]:.2f})  |  phd_get_reports_u8 n=1 pinned host {b[0]:.2f} ms "
          f"(min {b[1]:.2f})  |  device resident {c[0]:.2f} ms (min {c[1]:.2f})  stages " +
          " ".join(f"{k}={v:.3f}" for k, v in ms.items()), flush=True)


run("config1 1920x1080, no boxes", 1920, 1080, None)
W, H = 3840, 2160
run("config2 3840x2160 + 4 boxes", W, H,
    [dict(top=H * i // 8, bottom=H * i // 8 + H // 4, left=W * i // 8, right=W * i // 8 + W // 4) for i in range(4)])

# the Python front door: PIL/numpy image -> Report, through the byte route and through the reference's float64 planes
import photohive_dsp_b200 as P  # noqa: E402
from photohive_dsp_b200 import core  # noqa: E402

img = Generator(1920, 1080, dev).batch(1).cpu().numpy()[0]
for via in (False, True):
    core._VIA_DOUBLES = via
    for _ in range(2):
        P.get_report(img)
    m = med(lambda: P.get_report(img), 7)
    print(f"python get_report(1080p uint8) {'via float64 planes + C entry point' if via else 'byte route'}: {m[0]:.2f} ms (min {m[1]:.2f})", flush=True)
