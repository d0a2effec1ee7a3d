"""Batch front door: many same-size 8-bit RGB images per call, one context per GPU.

This is the additive interface the throughput metric is measured on (include/photohive_dsp.h, Part 2).
Inputs may be numpy uint8 arrays (host) or torch CUDA uint8 tensors (device); the reports come back as
numpy structured views over the flat records.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .lib import lib
from .structures import phd_flat_head, phd_flat_layout, phd_params

DEFAULTS = dict(h_partitions=18, s_partitions=2, v_partitions=3, black_thresh=0.1, gray_thresh=0.1,
                coverage_thresh=0.95, linked_list_size=1000, downsample_rate=1, radius_partitions=40,
                angle_partitions=72, quantity_weight=0.1, saturation_value_weight=0.9,
                fft_streak_thresh=1.20, magnitude_thresh=0.3, blur_cutoff_ratio_denom=2)

ERRORS = {1: "rejected by the reference's pre-checks", 2: "bad parameters", 3: "no CUDA device", 4: "CUDA error",
          5: "unsupported size", 6: "image is not 8-bit"}


class PhotoHiveError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}: {ERRORS.get(code, '?')}] {msg}")
        self.code = code


def make_params(**kw) -> phd_params:
    d = dict(DEFAULTS)
    unknown = set(kw) - set(d)
    if unknown:
        raise TypeError(f"unknown parameter(s): {sorted(unknown)}")
    d.update(kw)
    return phd_params(**d)


def flat_layout(params: phd_params, max_boxes: int = 0) -> phd_flat_layout:
    lay = phd_flat_layout()
    rc = lib.phd_flat_get_layout(C.byref(params), max_boxes, C.byref(lay))
    if rc != 0:
        raise PhotoHiveError(rc, "cannot lay out records for these parameters")
    return lay


@dataclass
class BatchReports:
    """Struct-of-arrays view of n flat records (numpy, host)."""
    raw: np.ndarray                 # [n, record_bytes] uint8
    layout: phd_flat_layout
    rgb_stats: np.ndarray           # [n, 6]
    average_saturation: np.ndarray  # [n]
    palette_n: np.ndarray           # [n]
    palette_hsv: np.ndarray         # [n, T, 3]  (first palette_n[i] rows valid)
    palette_pct: np.ndarray         # [n, T]
    parent_ids: np.ndarray          # [n, T]
    blur_bins: np.ndarray           # [n, na, nr]
    angle_bin_size: np.ndarray
    radius_bin_size: np.ndarray
    blur_vec_angle: np.ndarray      # [n, 10]
    blur_vec_mag: np.ndarray        # [n, 10]
    sharpness: np.ndarray | None    # [n, max_boxes]
    max_power: np.ndarray
    tie_groups: np.ndarray
    dropped_pixels: np.ndarray

    def __len__(self):
        return self.raw.shape[0]

    # ---- export: the reference's Report.to_json() schema (core.py:388-436), for a whole batch --------------
    def palette_rgb(self) -> np.ndarray:
        """[n, T, 3] integer RGB of every palette entry, vectorised twin of utils.hsv_to_rgb (utils.py:8-28)."""
        h, s, v = self.palette_hsv[..., 0], self.palette_hsv[..., 1], self.palette_hsv[..., 2]
        chroma = v * s
        second = chroma * (1 - np.abs((h / 60) % 2 - 1))
        base = v - chroma
        sector = np.minimum((h // 60).astype(np.int64), 5)
        zero = np.zeros_like(chroma)
        r = np.choose(sector, [chroma, second, zero, zero, second, chroma])
        g = np.choose(sector, [second, chroma, chroma, second, zero, zero])
        b = np.choose(sector, [zero, zero, second, chroma, chroma, second])
        with np.errstate(invalid="ignore"):
            return np.stack([((r + base) * 255), ((g + base) * 255), ((b + base) * 255)], -1).astype(np.int64)

    def to_dicts(self, height: int, width: int) -> list[dict]:
        """One dict per record with exactly the keys (and key order) of the reference's Report.to_json()."""
        rgb = self.palette_rgb()
        out = []
        names = ("Red Brightness", "Green Brightness", "Blue Brightness", "Red Contrast", "Green Contrast", "Blue Contrast")
        for i in range(len(self)):
            d = {"Height": int(height), "Width": int(width), "Average Saturation": float(self.average_saturation[i])}
            d.update({k: float(x) for k, x in zip(names, self.rgb_stats[i])})
            for k in range(10):
                d[f"Blur Vector {k+1} Angle"] = int(self.blur_vec_angle[i, k])
                d[f"Blur Vector {k+1} Magnitude"] = float(self.blur_vec_mag[i, k])
            n = int(self.palette_n[i])
            for k in range(100):
                if k < n:
                    c = rgb[i, k]
                    d[f"Color {k+1} H"], d[f"Color {k+1} S"], d[f"Color {k+1} V"] = int(c[0]), int(c[1]), int(c[2])
                    d[f"Color {k+1} Percentage"] = float(self.palette_pct[i, k])
                else:
                    d[f"Color {k+1} H"] = d[f"Color {k+1} S"] = d[f"Color {k+1} V"] = 0
                    d[f"Color {k+1} Percentage"] = 0
            ns = 0 if self.sharpness is None else self.sharpness.shape[1]
            for k in range(10):
                d[f"Sharpness {k+1}:"] = float(self.sharpness[i, k]) if k < ns else 0.0
            out.append(d)
        return out

    def to_json(self, i: int, height: int, width: int) -> str:
        """JSON text of record i, formatted like the reference's Report.to_json()."""
        import json
        one = BatchReports(**{k: (v[i:i + 1] if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        return json.dumps(one.to_dicts(height, width)[0], indent=4)


def view_records(raw: np.ndarray, lay: phd_flat_layout) -> BatchReports:
    n = raw.shape[0]
    T, na, nr, mb = lay.T, lay.na, lay.nr, lay.max_boxes
    head_dt = np.dtype([("rgb_stats", "<f8", 6), ("average_saturation", "<f8"), ("max_power", "<f8"),
                        ("dropped_pixels", "<i8"), ("palette_n", "<i4"), ("tie_groups", "<i4"),
                        ("n_sharpness", "<i4"), ("angle_bin_size", "<i4"), ("radius_bin_size", "<i4"),
                        ("num_angle_bins", "<i4"), ("num_radius_bins", "<i4"), ("status", "<i4"),
                        ("blur_vec_angle", "<i4", 10), ("blur_vec_mag", "<f4", 10)])
    assert head_dt.itemsize == C.sizeof(phd_flat_head)

    def field(off, dtype, count):
        dt = np.dtype(dtype)
        return np.ndarray((n, count), dt, raw, offset=off, strides=(raw.strides[0], dt.itemsize))

    head = np.ndarray((n,), head_dt, raw, offset=0, strides=(raw.strides[0],))
    bad = np.nonzero(head["status"])[0]
    if bad.size:
        raise PhotoHiveError(int(head["status"][bad[0]]), f"record {int(bad[0])} is invalid")
    return BatchReports(
        raw=raw, layout=lay, rgb_stats=head["rgb_stats"], average_saturation=head["average_saturation"],
        palette_n=head["palette_n"], palette_hsv=field(lay.off_palette_hsv, "<f8", 3 * T).reshape(n, T, 3),
        palette_pct=field(lay.off_palette_pct, "<f8", T), parent_ids=field(lay.off_parent_ids, "<i4", T),
        blur_bins=field(lay.off_blur_bins, "<f8", na * nr).reshape(n, na, nr),
        angle_bin_size=head["angle_bin_size"], radius_bin_size=head["radius_bin_size"],
        blur_vec_angle=head["blur_vec_angle"], blur_vec_mag=head["blur_vec_mag"],
        sharpness=field(lay.off_sharpness, "<f8", mb) if mb > 0 else None, max_power=head["max_power"],
        tie_groups=head["tie_groups"], dropped_pixels=head["dropped_pixels"])


class Context:
    """One per GPU: owns the stream, FFT/bin-map plans and workspaces of that device."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.phd_context_create(device, C.byref(h))
        if rc != 0:
            raise PhotoHiveError(rc, f"cannot create a context on CUDA device {device}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and lib is not None:  # `lib` is already gone when a module-level context dies at exit
            lib.phd_context_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise PhotoHiveError(rc, (lib.phd_last_error(self._h) or b"").decode(errors="replace"))

    # ---- the hot path -----------------------------------------------------------------------------
    def get_reports(self, images, boxes=None, params: phd_params | None = None, records_out=None,
                    **param_overrides) -> BatchReports:
        """images: uint8 [n,H,W,3] numpy array (host) or torch CUDA tensor (device, contiguous).
        boxes: None or int array [n, max_boxes, 4] of (top, bottom, left, right)."""
        params = params or make_params(**param_overrides)
        ptr, n, H, W, stride, keep = _as_image_batch(images, self.device)
        mb, bptr, bkeep = 0, None, None
        if boxes is not None:
            bkeep = np.ascontiguousarray(boxes, np.int32)
            if bkeep.ndim != 3 or bkeep.shape[0] != n or bkeep.shape[2] != 4:
                raise ValueError("boxes must have shape [n, max_boxes, 4]")
            mb = bkeep.shape[1]
            bptr = bkeep.ctypes.data_as(C.c_void_p) if mb > 0 else None
        lay = flat_layout(params, mb)
        raw = records_out if records_out is not None else np.empty((n, lay.record_bytes), np.uint8)
        self._check(lib.phd_get_reports_u8(self._h, ptr, n, W, H, stride, bptr, mb, C.byref(params),
                                           raw.ctypes.data_as(C.c_void_p)))
        return view_records(raw, lay)

    def get_reports_raw(self, rgb_ptr: int, n: int, width: int, height: int, stride: int, params: phd_params,
                        records_ptr: int, boxes_ptr: int | None = None, max_boxes: int = 0) -> None:
        """Pointer-level call for harnesses that manage their own (pinned / device) buffers."""
        self._check(lib.phd_get_reports_u8(self._h, rgb_ptr, n, width, height, stride, boxes_ptr, max_boxes,
                                           C.byref(params), records_ptr))

    def last_timing(self):
        ms = (C.c_float * 8)()
        launches = lib.phd_last_timing(self._h, C.byref(ms))
        names = ["total", "frontend", "palette_select", "palette_ties", "fft_rows", "fft_cols_blur",
                 "sharpness", "finalize"]
        return dict(zip(names, [float(x) for x in ms])), launches

    STAGES = ["total", "frontend", "palette_select", "palette_ties", "fft_rows", "fft_cols_blur", "sharpness", "finalize"]

    def last_stage_launches(self):
        n = (C.c_int * 8)()
        self._check(lib.phd_last_stage_launches(self._h, C.byref(n)))
        return dict(zip(self.STAGES, [int(x) for x in n]))

    def last_fused(self) -> bool:
        """True when the last call ran the front end and the row FFT as one launch (its time is under "frontend")."""
        return bool(lib.phd_last_fused(self._h))

    # ---- test hooks ---------------------------------------------------------------------------------
    def debug_group_sweep(self, params: phd_params, exact: bool = False) -> np.ndarray:
        """Group id of all 2^24 colours: product path, or (exact=True) the FP64 transcription."""
        out = np.empty(1 << 24, np.uint16)
        fn = lib.phd_debug_group_sweep_exact if exact else lib.phd_debug_group_sweep
        self._check(fn(self._h, C.byref(params), out.ctypes.data_as(C.c_void_p)))
        return out

    def debug_bin_map(self, width, height, nr=40, na=72):
        m = np.empty((height, width // 2 + 1), np.uint16)
        c = np.empty(na * nr, np.int32)
        self._check(lib.phd_debug_bin_map(self._h, width, height, nr, na, m.ctypes.data_as(C.c_void_p),
                                          c.ctypes.data_as(C.c_void_p)))
        return m, c.reshape(na, nr)

    def debug_power_spectrum(self, rgb: np.ndarray) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W, _ = rgb.shape
        out = np.empty((H, W // 2 + 1), np.float32)
        self._check(lib.phd_debug_power_spectrum(self._h, rgb.ctypes.data_as(C.c_void_p), W, H,
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def debug_group_counts(self, rgb: np.ndarray, params: phd_params) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W, _ = rgb.shape
        T = params.h_partitions * params.s_partitions * params.v_partitions + params.v_partitions + 1
        out = np.empty(T, np.int32)
        self._check(lib.phd_debug_group_counts(self._h, rgb.ctypes.data_as(C.c_void_p), W, H, C.byref(params),
                                               out.ctypes.data_as(C.c_void_p)))
        return out


class MultiContext:
    """The GPUs of one box behind ONE call (phd_get_reports_u8_multi): one context per device, the batch is split into
    contiguous ranges, every device writes its records into its range of one host array."""

    def __init__(self, devices):
        self.contexts = [Context(d) for d in devices]
        self._arr = (C.c_void_p * len(self.contexts))(*[c._h for c in self.contexts])

    def close(self):
        for c in self.contexts:
            c.close()

    def get_reports_raw(self, rgb_ptr: int, n: int, width: int, height: int, stride: int, params: phd_params,
                        records_ptr: int, boxes_ptr: int | None = None, max_boxes: int = 0) -> None:
        rc = lib.phd_get_reports_u8_multi(self._arr, len(self.contexts), rgb_ptr, n, width, height, stride, boxes_ptr,
                                          max_boxes, C.byref(params), records_ptr)
        if rc != 0:
            msgs = [(lib.phd_last_error(c._h) or b"").decode(errors="replace") for c in self.contexts]
            raise PhotoHiveError(rc, "; ".join(m for m in msgs if m))

    def get_reports(self, images: np.ndarray, boxes=None, params: phd_params | None = None, **param_overrides) -> BatchReports:
        """images: uint8 [n,H,W,3] numpy array in HOST memory (pinned or pageable)."""
        params = params or make_params(**param_overrides)
        if not isinstance(images, np.ndarray):
            raise TypeError("MultiContext takes host (numpy) batches; a device tensor belongs to one GPU")
        ptr, n, H, W, stride, keep = _as_image_batch(images)
        mb, bptr, bkeep = 0, None, None
        if boxes is not None:
            bkeep = np.ascontiguousarray(boxes, np.int32)
            mb = bkeep.shape[1]
            bptr = bkeep.ctypes.data_as(C.c_void_p) if mb > 0 else None
        lay = flat_layout(params, mb)
        raw = np.empty((n, lay.record_bytes), np.uint8)
        self.get_reports_raw(ptr, n, W, H, stride, params, raw.ctypes.data_as(C.c_void_p), bptr, mb)
        return view_records(raw, lay)


def _as_image_batch(images, device=None):
    """-> (pointer, n, H, W, stride_bytes, keepalive)"""
    if isinstance(images, np.ndarray):
        a = images
        if a.ndim == 3:
            a = a[None]
        if a.ndim != 4 or a.shape[3] != 3 or a.dtype != np.uint8:
            raise ValueError("images must be uint8 [n,H,W,3]")
        a = np.ascontiguousarray(a)
        n, H, W, _ = a.shape
        return a.ctypes.data_as(C.c_void_p), n, H, W, H * W * 3, a
    # torch tensor (imported lazily: torch is plumbing for device memory only)
    import torch
    if isinstance(images, torch.Tensor):
        t = images
        if t.dim() == 3:
            t = t[None]
        if t.dim() != 4 or t.shape[3] != 3 or t.dtype != torch.uint8:
            raise ValueError("images must be uint8 [n,H,W,3]")
        t = t.contiguous()
        n, H, W, _ = t.shape
        if t.is_cuda:
            # ordering contract of phd_get_reports_u8 (include/photohive_dsp.h): the library launches on its own
            # non-blocking stream, so whatever produced the tensor (torch's current stream, the DMA tail of a
            # pageable .cuda() copy) must have finished before the pointer is handed over
            if device is not None and t.device.index != device:
                raise ValueError(f"images live on cuda:{t.device.index} but the context belongs to cuda:{device}")
            torch.cuda.current_stream(t.device).synchronize()
        return C.c_void_p(t.data_ptr()), n, H, W, H * W * 3, t
    raise TypeError(f"unsupported image container {type(images)!r}")
