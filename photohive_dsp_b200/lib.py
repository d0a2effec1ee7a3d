"""Loads PhotoHive_DSP_lib/libreport_data.so (the CUDA build) and declares its C ABI for ctypes.

Same role and same relative library location as the reference's lib.py:20-37.  There is deliberately no
fallback: if the shared library is missing the import fails with instructions, and if no CUDA device is
present the entry points return NULL / PHD_E_NO_DEVICE and print why.
"""
import ctypes as C
import os

from .structures import (Blur_Profile, Crop_Boundaries, Full_Report_Data, Image_PGM, Image_RGB, phd_flat_layout,
                         phd_params)

directory = os.path.dirname(os.path.abspath(__file__))
lib_path = os.environ.get("PHD_LIB_PATH") or os.path.join(directory, "PhotoHive_DSP_lib", "libreport_data.so")  # override: A/B runs

if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing. Build it with `python -m photohive_dsp_b200.build` (needs nvcc; "
        "there is no CPU implementation to fall back to).")

lib = C.CDLL(lib_path)

# ---- drop-in entry points (src/interface.h:16-26, src/blur_profile.h:78) -----------------------
lib.get_full_report_data.restype = C.POINTER(Full_Report_Data)
lib.get_full_report_data.argtypes = [
    C.POINTER(Image_RGB), C.POINTER(Crop_Boundaries),
    C.c_int, C.c_int, C.c_int,            # h, s, v partitions
    C.c_double, C.c_double,               # black, gray thresholds
    C.c_double, C.c_int,                  # coverage threshold, linked list size
    C.c_int, C.c_int, C.c_int,            # downsample rate, radius, angle partitions
    C.c_float, C.c_float,                 # quantity weight, saturation-value weight
    C.c_double, C.c_double, C.c_int,      # streak threshold, magnitude threshold, cutoff denominator
]
lib.free_full_report.restype = None
lib.free_full_report.argtypes = [C.POINTER(C.POINTER(Full_Report_Data))]
lib.get_blur_profile_visual.restype = C.POINTER(Image_PGM)
lib.get_blur_profile_visual.argtypes = [C.POINTER(Blur_Profile), C.c_int, C.c_int]

# ---- batch interface -----------------------------------------------------------------------------
lib.phd_default_params.restype = None
lib.phd_default_params.argtypes = [C.POINTER(phd_params)]
lib.phd_context_create.restype = C.c_int
lib.phd_context_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
lib.phd_context_destroy.restype = None
lib.phd_context_destroy.argtypes = [C.c_void_p]
lib.phd_last_error.restype = C.c_char_p
lib.phd_last_error.argtypes = [C.c_void_p]
lib.phd_flat_get_layout.restype = C.c_int
lib.phd_flat_get_layout.argtypes = [C.POINTER(phd_params), C.c_int, C.POINTER(phd_flat_layout)]
lib.phd_get_reports_u8.restype = C.c_int
lib.phd_get_reports_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                   C.c_int, C.POINTER(phd_params), C.c_void_p]
lib.phd_get_reports_u8_multi.restype = C.c_int
lib.phd_get_reports_u8_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                         C.c_void_p, C.c_int, C.POINTER(phd_params), C.c_void_p]
lib.phd_flat_to_full_report.restype = C.POINTER(Full_Report_Data)
lib.phd_flat_to_full_report.argtypes = [C.c_void_p, C.POINTER(phd_flat_layout)]
lib.phd_last_timing.restype = C.c_int
lib.phd_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float * 8)]
lib.phd_last_stage_launches.restype = C.c_int
lib.phd_last_stage_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int * 8)]
lib.phd_last_fused.restype = C.c_int
lib.phd_last_fused.argtypes = [C.c_void_p]
lib.phd_debug_group_sweep.restype = C.c_int
lib.phd_debug_group_sweep.argtypes = [C.c_void_p, C.POINTER(phd_params), C.c_void_p]
lib.phd_debug_group_sweep_exact.restype = C.c_int
lib.phd_debug_group_sweep_exact.argtypes = [C.c_void_p, C.POINTER(phd_params), C.c_void_p]
lib.phd_debug_bin_map.restype = C.c_int
lib.phd_debug_bin_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.phd_debug_power_spectrum.restype = C.c_int
lib.phd_debug_power_spectrum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.phd_debug_group_counts.restype = C.c_int
lib.phd_debug_group_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(phd_params), C.c_void_p]

EXPORTED = ["get_full_report_data", "free_full_report", "get_blur_profile_visual", "phd_default_params",
            "phd_context_create", "phd_context_destroy", "phd_last_error", "phd_flat_get_layout",
            "phd_get_reports_u8", "phd_get_reports_u8_multi", "phd_flat_to_full_report", "phd_last_timing", "phd_last_stage_launches", "phd_last_fused",
            "phd_debug_group_sweep",
            "phd_debug_group_sweep_exact",
            "phd_debug_bin_map", "phd_debug_power_spectrum", "phd_debug_group_counts"]
