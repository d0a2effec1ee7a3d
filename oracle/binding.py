"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle and the compiled reference.

* ``Oracle``    : oracle/libphd_oracle.so, our CPU restatement (photohive_oracle.c).
* ``Reference`` : oracle/_ref/libreport_data_ref_O{0,2}.so, the UNMODIFIED reference sources compiled by
                  ``make -C oracle ref`` against the FFTW stand-in (only possible where /root/reference exists;
                  the GPU box uses the prebuilt file shipped by gpurun).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (photohive_dsp_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libphd_oracle.so")
REF_SO = {0: os.path.join(HERE, "_ref", "libreport_data_ref_O0.so"),
          2: os.path.join(HERE, "_ref", "libreport_data_ref_O2.so")}

DEFAULTS = dict(h_partitions=18, s_partitions=2, v_partitions=3, black_thresh=0.1, gray_thresh=0.1,
                coverage_thresh=0.95, linked_list_size=1000, downsample_rate=1, radius_partitions=40,
                angle_partitions=72, quantity_weight=0.1, saturation_value_weight=0.9,
                fft_streak_thresh=1.20, magnitude_thresh=0.3, blur_cutoff_ratio_denom=2)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference (when its sources are present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class Params(C.Structure):
    _fields_ = [("h_partitions", C.c_int), ("s_partitions", C.c_int), ("v_partitions", C.c_int),
                ("black_thresh", C.c_double), ("gray_thresh", C.c_double), ("coverage_thresh", C.c_double),
                ("linked_list_size", C.c_int), ("downsample_rate", C.c_int), ("radius_partitions", C.c_int),
                ("angle_partitions", C.c_int), ("quantity_weight", C.c_float),
                ("saturation_value_weight", C.c_float), ("fft_streak_thresh", C.c_double),
                ("magnitude_thresh", C.c_double), ("blur_cutoff_ratio_denom", C.c_int)]


def make_params(**kw) -> Params:
    d = dict(DEFAULTS)
    d.update(kw)
    return Params(**d)


class _Out(C.Structure):
    _fields_ = [("rgb_stats", C.c_double * 6), ("average_saturation", C.c_double),
                ("hsv_width", C.c_int), ("hsv_height", C.c_int), ("T", C.c_int),
                ("group_counts", C.c_void_p), ("group_ids", C.c_void_p), ("gray", C.c_void_p),
                ("palette_n", C.c_int), ("parent_ids", C.c_void_p), ("palette_hsv", C.c_void_p),
                ("palette_pct", C.c_void_p), ("tie_groups", C.c_int), ("dropped_pixels", C.c_long),
                ("angle_bin_size", C.c_int), ("radius_bin_size", C.c_int),
                ("blur_bins", C.c_void_p), ("blur_counts", C.c_void_p), ("bin_map", C.c_void_p),
                ("power", C.c_void_p), ("max_power", C.c_double),
                ("blur_vec_angle", C.c_int * 10), ("blur_vec_mag", C.c_float * 10),
                ("sharpness", C.c_void_p)]


@dataclass
class Report:
    """Plain-numpy view of one report, common to oracle, reference and product."""
    rgb_stats: np.ndarray = None          # [6]
    average_saturation: float = 0.0
    palette_hsv: np.ndarray = None        # [N,3]
    palette_pct: np.ndarray = None        # [N]
    blur_bins: np.ndarray = None          # [na,nr]
    angle_bin_size: int = 0
    radius_bin_size: int = 0
    blur_vec_angle: np.ndarray = None     # [10] int
    blur_vec_mag: np.ndarray = None       # [10] float32
    sharpness: np.ndarray = None          # [nboxes] or None
    extra: dict = field(default_factory=dict)


def group_total(p: Params) -> int:
    return p.h_partitions * p.s_partitions * p.v_partitions + p.v_partitions + 1


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        assert self.lib.phd_oracle_sizeof_params() == C.sizeof(Params)
        assert self.lib.phd_oracle_sizeof_out() == C.sizeof(_Out)
        self.lib.phd_oracle_generate.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        self.lib.phd_oracle_generate.restype = None
        self.lib.phd_oracle_rejects.argtypes = [C.c_int, C.c_int]
        self.lib.phd_oracle_group_sweep.argtypes = [C.POINTER(Params), C.c_void_p]
        self.lib.phd_oracle_group_sweep.restype = None
        self.lib.phd_oracle_bin_map.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.phd_oracle_bin_map.restype = None
        self.lib.phd_oracle_hsv_of.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.lib.phd_oracle_hsv_of.restype = None
        for name in ("phd_oracle_report_u8",):
            f = getattr(self.lib, name)
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Params), C.c_int, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(_Out)]
            f.restype = C.c_int
        self.lib.phd_oracle_report_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                   C.POINTER(Params), C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(_Out)]
        self.lib.phd_oracle_report_f64.restype = C.c_int

    def generate(self, kind: int, seed: int, width: int, height: int) -> np.ndarray:
        out = np.empty((height, width, 3), np.uint8)
        self.lib.phd_oracle_generate(kind, seed, width, height, _ptr(out))
        return out

    def rejects(self, width: int, height: int) -> bool:
        return bool(self.lib.phd_oracle_rejects(width, height))

    def group_sweep(self, params: Params) -> np.ndarray:
        out = np.empty(1 << 24, np.uint16)
        self.lib.phd_oracle_group_sweep(C.byref(params), _ptr(out))
        return out

    def hsv_of(self, r: int, g: int, b: int) -> np.ndarray:
        out = np.empty(3, np.float64)
        self.lib.phd_oracle_hsv_of(r, g, b, _ptr(out))
        return out

    def bin_map(self, width, height, nr=40, na=72):
        fw = width // 2 + 1
        m = np.empty((height, fw), np.uint16)
        c = np.empty(na * nr, np.int32)
        self.lib.phd_oracle_bin_map(width, height, nr, na, _ptr(m), _ptr(c))
        return m, c.reshape(na, nr)

    def report(self, rgb: np.ndarray, params: Params | None = None, boxes=None, stages: int = 7,
               nthreads: int = 1, want_intermediates: bool = False, planes=None) -> Report | None:
        """rgb: uint8 [H,W,3] (or planes=(r,g,b) float64 [H,W] for non-8-bit inputs)."""
        params = params or make_params()
        if planes is not None:
            H, W = planes[0].shape
        else:
            rgb = np.ascontiguousarray(rgb, np.uint8)
            H, W, _ = rgb.shape
        T = group_total(params)
        na, nr = params.angle_partitions, params.radius_partitions
        fw = W // 2 + 1
        nb = 0
        tb = [None] * 4
        if boxes is not None:
            nb = len(boxes)
            tb = [np.array([b[k] for b in boxes], np.int32) for k in ("top", "bottom", "left", "right")]
        o = _Out()
        counts = np.zeros(T, np.int32)
        parent_ids = np.zeros(T, np.int32)
        pal_hsv = np.zeros(3 * T, np.float64)
        pal_pct = np.zeros(T, np.float64)
        bins = np.zeros(na * nr, np.float64)
        sharp = np.zeros(max(nb, 1), np.float64)
        o.group_counts, o.parent_ids = _ptr(counts), _ptr(parent_ids)
        o.palette_hsv, o.palette_pct = _ptr(pal_hsv), _ptr(pal_pct)
        o.blur_bins, o.sharpness = _ptr(bins), _ptr(sharp)
        keep = {}
        if want_intermediates:
            ds = max(params.downsample_rate, 1)
            hw, hh = (W // ds, H // ds) if ds > 1 else (W, H)
            keep["group_ids"] = np.zeros(hw * hh, np.uint16)
            keep["gray"] = np.zeros(W * H, np.float64)
            keep["blur_counts"] = np.zeros(na * nr, np.int32)
            keep["bin_map"] = np.zeros(H * fw, np.uint16)
            keep["power"] = np.zeros(H * fw, np.float64)
            o.group_ids, o.gray = _ptr(keep["group_ids"]), _ptr(keep["gray"])
            o.blur_counts, o.bin_map, o.power = _ptr(keep["blur_counts"]), _ptr(keep["bin_map"]), _ptr(keep["power"])
        if planes is not None:
            pr, pg, pb = (np.ascontiguousarray(p, np.float64) for p in planes)
            rc = self.lib.phd_oracle_report_f64(_ptr(pr), _ptr(pg), _ptr(pb), W, H, C.byref(params), nb,
                                                _ptr(tb[0]), _ptr(tb[1]), _ptr(tb[2]), _ptr(tb[3]), stages,
                                                nthreads, C.byref(o))
        else:
            rc = self.lib.phd_oracle_report_u8(_ptr(rgb), W, H, C.byref(params), nb, _ptr(tb[0]), _ptr(tb[1]),
                                               _ptr(tb[2]), _ptr(tb[3]), stages, nthreads, C.byref(o))
        if rc != 0:
            return None
        n = o.palette_n
        rep = Report(rgb_stats=np.array(o.rgb_stats[:]), average_saturation=o.average_saturation,
                     palette_hsv=pal_hsv[:3 * n].reshape(n, 3).copy(), palette_pct=pal_pct[:n].copy(),
                     blur_bins=bins.reshape(na, nr), angle_bin_size=o.angle_bin_size,
                     radius_bin_size=o.radius_bin_size, blur_vec_angle=np.array(o.blur_vec_angle[:], np.int32),
                     blur_vec_mag=np.array(o.blur_vec_mag[:], np.float32),
                     sharpness=sharp[:nb].copy() if boxes is not None else None)
        rep.extra = dict(group_counts=counts, parent_ids=parent_ids[:n].copy(), tie_groups=o.tie_groups,
                         dropped_pixels=o.dropped_pixels, max_power=o.max_power, T=T,
                         hsv_size=(o.hsv_width, o.hsv_height), **keep)
        return rep


# ---------------------------------------------------------------------------------------------------------------
# The compiled reference, called through its own C entry points (src/interface.h:16-26) with ctypes mirrors of
# its structs (layout per SURVEY.md section 8b).
# ---------------------------------------------------------------------------------------------------------------
class R_Image_RGB(C.Structure):
    _fields_ = [("height", C.c_int), ("width", C.c_int), ("r", C.POINTER(C.c_double)),
                ("g", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double))]


class R_Pixel_HSV(C.Structure):
    _fields_ = [("parent_id", C.c_int), ("h", C.c_double), ("s", C.c_double), ("v", C.c_double)]


class R_Crop(C.Structure):
    _fields_ = [("N", C.c_int), ("top", C.POINTER(C.c_int)), ("bottom", C.POINTER(C.c_int)),
                ("left", C.POINTER(C.c_int)), ("right", C.POINTER(C.c_int))]


class R_Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("Br", "Bg", "Bb", "Cr", "Cg", "Cb")]


class R_Palette(C.Structure):
    _fields_ = [("N", C.c_int), ("averages", C.POINTER(R_Pixel_HSV)), ("percentages", C.POINTER(C.c_double))]


class R_BlurProfile(C.Structure):
    _fields_ = [("num_angle_bins", C.c_int), ("num_radius_bins", C.c_int), ("angle_bin_size", C.c_int),
                ("radius_bin_size", C.c_int), ("bins", C.POINTER(C.POINTER(C.c_double)))]


class R_BlurVector(C.Structure):
    _fields_ = [("angle", C.c_int), ("magnitude", C.c_float)]


class R_BlurVectorGroup(C.Structure):
    _fields_ = [("len_vectors", C.c_int), ("blur_vectors", C.POINTER(R_BlurVector))]


class R_Sharpnesses(C.Structure):
    _fields_ = [("N", C.c_int), ("sharpness", C.POINTER(C.c_double))]


class R_Full(C.Structure):
    _fields_ = [("rgb_stats", C.POINTER(R_Stats)), ("color_palette", C.POINTER(R_Palette)),
                ("blur_profile", C.POINTER(R_BlurProfile)), ("blur_vectors", C.POINTER(R_BlurVectorGroup)),
                ("average_saturation", C.c_double), ("sharpness", C.POINTER(R_Sharpnesses))]


def unpack_full_report(rp, nboxes_expected=None) -> Report:
    """Copy a Full_Report_Data* (from the reference OR from the product's drop-in library) into numpy."""
    r = rp.contents
    st = r.rgb_stats.contents
    pal = r.color_palette.contents
    n = pal.N
    hsv = np.array([[pal.averages[i].h, pal.averages[i].s, pal.averages[i].v] for i in range(n)], np.float64).reshape(n, 3)
    pct = np.array([pal.percentages[i] for i in range(n)], np.float64)
    bp = r.blur_profile.contents
    na, nr = bp.num_angle_bins, bp.num_radius_bins
    bins = np.array([[bp.bins[a][j] for j in range(nr)] for a in range(na)], np.float64)
    bv = r.blur_vectors.contents
    ang = np.array([bv.blur_vectors[i].angle for i in range(bv.len_vectors)], np.int32)
    mag = np.array([bv.blur_vectors[i].magnitude for i in range(bv.len_vectors)], np.float32)
    sharp = None
    if r.sharpness:
        s = r.sharpness.contents
        sharp = np.array([s.sharpness[i] for i in range(s.N)], np.float64)
    return Report(rgb_stats=np.array([st.Br, st.Bg, st.Bb, st.Cr, st.Cg, st.Cb]),
                  average_saturation=r.average_saturation, palette_hsv=hsv, palette_pct=pct, blur_bins=bins,
                  angle_bin_size=bp.angle_bin_size, radius_bin_size=bp.radius_bin_size, blur_vec_angle=ang,
                  blur_vec_mag=mag, sharpness=sharp, extra=dict(len_vectors=bv.len_vectors))


def planes_from_u8(rgb: np.ndarray):
    """utils.py:30-37: np.array(img)/255.0 -> three contiguous float64 planes."""
    a = rgb.astype(np.float64) / 255.0
    return (np.ascontiguousarray(a[:, :, 0]).ravel(), np.ascontiguousarray(a[:, :, 1]).ravel(),
            np.ascontiguousarray(a[:, :, 2]).ravel())


def make_crop(boxes):
    n = len(boxes)
    arrs = [(C.c_int * n)(*[int(b[k]) for b in boxes]) for k in ("top", "bottom", "left", "right")]
    cb = R_Crop(N=n, top=arrs[0], bottom=arrs[1], left=arrs[2], right=arrs[3])
    cb._keep = arrs
    return cb


def bind_entry_points(lib):
    lib.get_full_report_data.restype = C.POINTER(R_Full)
    lib.get_full_report_data.argtypes = [C.POINTER(R_Image_RGB), C.POINTER(R_Crop), C.c_int, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_double, C.c_double, C.c_int]
    lib.free_full_report.restype = None
    lib.free_full_report.argtypes = [C.POINTER(C.POINTER(R_Full))]
    return lib


def call_entry_point(lib, planes, width, height, params: Params, boxes=None):
    """Call get_full_report_data the way core.py:442-469 does; returns the raw pointer (may be NULL)."""
    r, g, b = planes
    img = R_Image_RGB(height=height, width=width, r=r.ctypes.data_as(C.POINTER(C.c_double)),
                      g=g.ctypes.data_as(C.POINTER(C.c_double)), b=b.ctypes.data_as(C.POINTER(C.c_double)))
    cb = make_crop(boxes) if boxes is not None else None
    cbp = C.byref(cb) if cb is not None else C.POINTER(R_Crop)()
    p = params
    return lib.get_full_report_data(C.byref(img), cbp, p.h_partitions, p.s_partitions, p.v_partitions,
                                    p.black_thresh, p.gray_thresh, p.coverage_thresh, p.linked_list_size,
                                    p.downsample_rate, p.radius_partitions, p.angle_partitions,
                                    p.quantity_weight, p.saturation_value_weight, p.fft_streak_thresh,
                                    p.magnitude_thresh, p.blur_cutoff_ratio_denom)


class Reference:
    """The unmodified reference, compiled into oracle/_ref/ (opt=0 parity build, opt=2 speed build)."""

    def __init__(self, opt: int = 0):
        path = REF_SO[opt]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = bind_entry_points(C.CDLL(path))

    @staticmethod
    def available(opt: int = 0) -> bool:
        return os.path.exists(REF_SO[opt])

    def report(self, rgb: np.ndarray, params: Params | None = None, boxes=None, planes=None) -> Report | None:
        params = params or make_params()
        if planes is None:
            H, W, _ = rgb.shape
            planes = planes_from_u8(rgb)
        else:
            H, W = planes[0].shape
            planes = tuple(np.ascontiguousarray(p, np.float64).ravel() for p in planes)
        rp = call_entry_point(self.lib, planes, W, H, params, boxes)
        if not rp:
            return None
        rep = unpack_full_report(rp)
        self.lib.free_full_report(C.byref(rp))
        return rep
