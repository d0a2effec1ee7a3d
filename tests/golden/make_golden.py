"""Generates tests/golden/reference_golden.npz from the UNMODIFIED reference.

Run it where /root/reference exists:  python tests/golden/make_golden.py
It builds oracle/_ref/libreport_data_ref_O0.so (the reference's own sources at -O0 against the FFTW
stand-in, oracle/Makefile `ref`), calls its get_full_report_data on seeded synthetic images and stores every
report field.  Images are not stored: they are regenerated from (kind, seed, width, height) by the oracle's
generator and guarded by a CRC32, so a generator change cannot silently re-define the fixtures.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding  # noqa: E402
from parity import planes_for  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")


def boxes_for(W, H, n=4):
    return [dict(top=H * i // 8, bottom=H * i // 8 + H // 4, left=W * i // 8, right=W * i // 8 + W // 4) for i in range(n)]


# name -> (kind, seed, W, H, params overrides, with boxes?)
CASES = {
    "g1_small": (1, 12345, 480, 360, {}, True),
    "g0_small": (0, 12346, 480, 360, {}, True),
    "g2_small": (2, 12347, 480, 360, {}, False),
    "g1_odd": (1, 7, 401, 357, {}, True),                       # odd width and height (bin-map overwrite rule)
    "g1_odd2": (1, 8, 405, 357, {}, True),                      # odd sizes whose factors (3,5,7,17) the GPU FFT covers
    "g0_fine": (0, 99, 400, 400, dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99), False),
    "g1_down5": (1, 5, 640, 480, dict(downsample_rate=5), True),
    "g1_list50": (1, 11, 480, 360, dict(linked_list_size=50), False),   # small list nodes: tie path drops pixels
    "g2_cov1": (2, 3, 480, 360, dict(coverage_thresh=1.0), False),
    "g1_1080p": (1, 12345, 1920, 1080, {}, True),               # SURVEY.md B.3 known answers
    "g0_1080p": (0, 12345, 1920, 1080, {}, False),
    "g2_1080p": (2, 12345, 1920, 1080, {}, False),
}


# name -> (plane kind, seed, W, H, params overrides, with boxes?)
PLANE_CASES = {
    "f64_test_rgb": ("test_rgb", 0, 400, 360, {}, False),
    "f64_random": ("random", 11, 480, 360, {}, True),
    "f64_random_ties": ("random", 12, 480, 360, dict(linked_list_size=40, coverage_thresh=0.6, downsample_rate=2), True),
    "f64_k65535": ("k65535", 13, 401, 357, dict(h_partitions=12, s_partitions=3, v_partitions=4), False),
}


def store_report(store, name, r):
    store[f"{name}/rgb_stats"] = r.rgb_stats
    store[f"{name}/average_saturation"] = np.array(r.average_saturation)
    store[f"{name}/palette_hsv"] = r.palette_hsv
    store[f"{name}/palette_pct"] = r.palette_pct
    store[f"{name}/blur_bins"] = r.blur_bins
    store[f"{name}/ints"] = np.array([r.angle_bin_size, r.radius_bin_size], np.int32)
    store[f"{name}/blur_vec_angle"] = r.blur_vec_angle
    store[f"{name}/blur_vec_mag"] = r.blur_vec_mag
    if r.sharpness is not None:
        store[f"{name}/sharpness"] = r.sharpness


def main():
    binding.build(ref=True)
    orc = binding.Oracle()
    ref = binding.Reference(0)
    store, meta = {}, {}
    for name, (kind, seed, W, H, kw, with_boxes) in PLANE_CASES.items():
        planes = planes_for(kind, seed, W, H)
        bx = boxes_for(W, H) if with_boxes else None
        r = ref.report(None, binding.make_params(**kw), boxes=bx, planes=planes)
        assert r is not None, name
        crc = zlib.crc32(np.stack(planes).tobytes())
        meta[name] = dict(planes=kind, seed=seed, W=W, H=H, params=kw, boxes=bx, crc32=crc)
        store_report(store, name, r)
        print(name, "N=%d sum%%=%.12f" % (len(r.palette_pct), r.palette_pct.sum()))
    for name, (kind, seed, W, H, kw, with_boxes) in CASES.items():
        img = orc.generate(kind, seed, W, H)
        bx = boxes_for(W, H) if with_boxes else None
        r = ref.report(img, binding.make_params(**kw), boxes=bx)
        assert r is not None, name
        meta[name] = dict(kind=kind, seed=seed, W=W, H=H, params=kw, boxes=bx, crc32=zlib.crc32(img.tobytes()))
        store_report(store, name, r)
        print(name, "N=%d sum%%=%.12f" % (len(r.palette_pct), r.palette_pct.sum()))
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
