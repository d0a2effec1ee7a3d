"""The drop-in boundary without a GPU: layout, exports, ownership, rejection rules, loud failure."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "photohive_dsp.h")


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def phd():
    from photohive_dsp_b200 import lib as L
    return L


def test_struct_sizes_match_reference_abi():
    """SURVEY.md section 8b (x86-64 SysV sizes of the reference structs)."""
    from photohive_dsp_b200 import structures as S
    for name, size in S.ABI_SIZES.items():
        assert C.sizeof(getattr(S, name)) == size, name
    assert S.Image_RGB.height.offset == 0 and S.Image_RGB.width.offset == 4 and S.Image_RGB.r.offset == 8
    assert S.Pixel_HSV.h.offset == 8 and S.Crop_Boundaries.top.offset == 8 and S.Blur_Profile.bins.offset == 16
    assert S.Full_Report_Data.average_saturation.offset == 32 and S.Full_Report_Data.sharpness.offset == 40


def test_header_layout_matches_ctypes(tmp_path):
    """Compile the public header with gcc and compare sizeof/offsetof with the ctypes mirrors."""
    from photohive_dsp_b200 import structures as S
    names = list(S.ABI_SIZES) + ["phd_params", "phd_flat_head", "phd_flat_layout"]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for n in names:
        lines.append(f'printf("{n} %zu\\n", sizeof({n}));')
    lines += ['printf("off_avg_sat %zu\\n", offsetof(Full_Report_Data, average_saturation));',
              'printf("off_vec_mag %zu\\n", offsetof(phd_flat_head, blur_vec_mag));', "return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-std=c11", str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(S, n)), n
    assert int(out["off_avg_sat"]) == 32
    assert int(out["off_vec_mag"]) == S.phd_flat_head.blur_vec_mag.offset


def test_library_exports_every_declared_symbol(phd):
    declared = re.findall(r"PHD_API\s+[\w\s\*]+?\b(\w+)\s*\(", open(HEADER).read())
    assert set(declared) == set(phd.EXPORTED)
    for name in declared:
        assert hasattr(phd.lib, name), f"{name} is declared in include/photohive_dsp.h but not exported"
    assert {"get_full_report_data", "free_full_report", "get_blur_profile_visual"} <= set(declared)


def test_rejection_rules_through_the_c_abi(phd, capfd):
    """The reference's own three rejection tests (src/test/test.c:87-135): NULL before any GPU work."""
    from photohive_dsp_b200.structures import Crop_Boundaries, Image_RGB
    dp = C.POINTER(C.c_double)
    plane = np.zeros(2001 * 400, np.float64)
    p = plane.ctypes.data_as(dp)
    args = (18, 2, 3, 0.1, 0.1, 0.95, 1000, 1, 40, 72, 0.1, 0.9, 1.15, 0.3, 2)
    null_boxes = C.POINTER(Crop_Boundaries)()
    for (w, h) in [(120000, 10000), (2001, 400), (400, 2001), (349, 350)]:
        img = Image_RGB(height=h, width=w, r=p, g=p, b=p)
        assert not phd.lib.get_full_report_data(C.byref(img), null_boxes, *args)
    img = Image_RGB(height=400, width=400, r=p, g=None, b=p)
    assert not phd.lib.get_full_report_data(C.byref(img), null_boxes, *args)
    assert not phd.lib.get_full_report_data(None, null_boxes, *args)
    err = capfd.readouterr().err
    assert "Invalid aspect ratio" in err and "greater than 350" in err and "NULL" in err


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly_not_silently(phd, capfd):
    """No CPU fallback: without a CUDA device the context cannot be made and a valid image yields NULL."""
    from photohive_dsp_b200.batch import Context, PhotoHiveError
    from photohive_dsp_b200.structures import Crop_Boundaries, Image_RGB
    with pytest.raises(PhotoHiveError) as e:
        Context(0)
    assert e.value.code == 3
    plane = np.full(400 * 400, 0.5019607843137255, np.float64)
    p = plane.ctypes.data_as(C.POINTER(C.c_double))
    img = Image_RGB(height=400, width=400, r=p, g=p, b=p)
    r = phd.lib.get_full_report_data(C.byref(img), C.POINTER(Crop_Boundaries)(), 18, 2, 3, 0.1, 0.1, 0.95, 1000, 1, 40,
                                     72, 0.1, 0.9, 1.2, 0.3, 2)
    assert not r
    assert "no CUDA device" in capfd.readouterr().err


def _fake_record(phd, max_boxes):
    from photohive_dsp_b200.batch import flat_layout, make_params
    from photohive_dsp_b200.structures import phd_flat_head
    p = make_params()
    lay = flat_layout(p, max_boxes)
    raw = np.zeros(lay.record_bytes, np.uint8)
    head = phd_flat_head.from_buffer(raw)
    for i in range(6):
        head.rgb_stats[i] = 0.1 * (i + 1)
    head.average_saturation = 0.25
    head.palette_n = 3
    head.n_sharpness = max_boxes if max_boxes > 0 else -1
    head.num_angle_bins, head.num_radius_bins = lay.na, lay.nr
    head.angle_bin_size, head.radius_bin_size = 2, 27
    head.blur_vec_angle[0], head.blur_vec_mag[0] = 2, 0.15
    hsv = np.ndarray((lay.T, 3), np.float64, raw, lay.off_palette_hsv)
    hsv[:3] = [[10, .5, .6], [200, .7, .8], [0, 0, .3]]
    np.ndarray((lay.T,), np.float64, raw, lay.off_palette_pct)[:3] = [.5, .3, .2]
    np.ndarray((lay.T,), np.int32, raw, lay.off_parent_ids)[:3] = [4, 60, 109]
    np.ndarray((lay.na, lay.nr), np.float64, raw, lay.off_blur_bins)[:] = np.arange(lay.na * lay.nr).reshape(lay.na, lay.nr)
    if max_boxes:
        np.ndarray((max_boxes,), np.float64, raw, lay.off_sharpness)[:] = [1.5, 2.5][:max_boxes]
    return raw, lay


@pytest.mark.parametrize("max_boxes", [0, 2])
def test_full_report_assembly_and_ownership(phd, max_boxes):
    """compile_full_report / free_full_report contract (src/utilities.c:210-226, src/interface.c:97-111)."""
    from photohive_dsp_b200.core import Report
    raw, lay = _fake_record(phd, max_boxes)
    ptr = phd.lib.phd_flat_to_full_report(raw.ctypes.data_as(C.c_void_p), C.byref(lay))
    assert ptr
    r = ptr.contents
    assert r.rgb_stats.contents.Cb == pytest.approx(0.6) and r.average_saturation == 0.25
    assert r.color_palette.contents.N == 3 and r.color_palette.contents.averages[1].h == 200
    assert r.color_palette.contents.averages[2].parent_id == 109
    assert r.blur_profile.contents.bins[71][39] == 72 * 40 - 1
    assert r.blur_vectors.contents.len_vectors == 10          # core.py:406-410 indexes all ten
    assert bool(r.sharpness) == (max_boxes > 0)
    rep = Report(ptr, 1080, 1920)
    js = rep.to_json()
    assert '"Blur Vector 1 Angle": 2' in js and '"Color 100 Percentage": 0' in js and '"Sharpness 10:": 0.0' in js
    assert rep.sharpnesses == ([1.5, 2.5] if max_boxes else [])
    assert rep.color_palette.colors[0] == (153, 89, 76)
    del rep  # frees through free_full_report
    ptr2 = phd.lib.phd_flat_to_full_report(raw.ctypes.data_as(C.c_void_p), C.byref(lay))
    phd.lib.free_full_report(C.byref(ptr2))
    assert not ptr2  # the caller's pointer is NULLed (interface.c:109)


def test_blur_profile_visual_matches_restatement(phd):
    """get_blur_profile_visual (src/blur_profile.c:140-180) is host code: checkable here."""
    raw, lay = _fake_record(phd, 0)
    ptr = phd.lib.phd_flat_to_full_report(raw.ctypes.data_as(C.c_void_p), C.byref(lay))
    bp = ptr.contents.blur_profile
    H, W = 90, 120
    pgm = phd.lib.get_blur_profile_visual(bp, H, W)
    got = np.ctypeslib.as_array(pgm.contents.data, shape=(H * W,)).reshape(H, W).copy()
    bins = np.arange(72 * 40, dtype=np.float64).reshape(72, 40)
    y, x = np.mgrid[0:H, 0:W]
    dy = np.where(y < H // 2, -y, H - y).astype(np.float64)
    r = np.sqrt(x.astype(np.float64) ** 2 + dy ** 2)
    rb = np.minimum((r / 27).astype(np.int64), 39)
    pb = ((np.arctan2(dy, x.astype(np.float64)) + 3.14159265 * 0.5) / 3.14159265 * 71.0).astype(np.int64)
    pb = np.clip(pb, 0, 71)
    assert np.array_equal(got, bins[pb, rb])
    C.CDLL(None).free(C.cast(pgm.contents.data, C.c_void_p))
    phd.lib.free_full_report(C.byref(ptr))


def test_layout_is_consistent(phd):
    from photohive_dsp_b200.batch import flat_layout, make_params
    for kw, mb in [({}, 0), ({}, 4), (dict(h_partitions=36, s_partitions=4, v_partitions=6), 10)]:
        lay = flat_layout(make_params(**kw), mb)
        assert lay.record_bytes % 16 == 0
        offs = [lay.off_palette_hsv, lay.off_palette_pct, lay.off_parent_ids, lay.off_blur_bins, lay.off_sharpness]
        assert offs == sorted(offs) and all(o % 8 == 0 for o in offs)
        assert lay.off_sharpness + 8 * mb <= lay.record_bytes


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under photohive_dsp_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "photohive_dsp_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"import\s+oracle|from\s+oracle|libphd_oracle|oracle/|dfft\.h|fftw3\.h|cufft", text), \
                    f"{f} refers to the oracle / a library FFT"
    lib = os.path.join(pkg, "PhotoHive_DSP_lib", "libreport_data.so")
    syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
    assert "phd_oracle" not in syms and "dfft_" not in syms and "fftw_" not in syms
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "cufft" not in needed.lower()
