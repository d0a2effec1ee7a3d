"""ctypes mirrors of the structs in include/photohive_dsp.h.

Part 1 types keep the names and field order of the reference's ``structures.py`` (they describe the same C
ABI: src/image_processing.h:12-98, src/color_quantization.h:11-15, src/blur_profile.h:20-61,
src/utilities.h:25-37).  Part 2 types belong to the additive batch interface.
"""
import ctypes as C

Pixel = C.c_double
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _struct(name, fields, methods=None):
    ns = {"_fields_": fields}
    ns.update(methods or {})
    return type(name, (C.Structure,), ns)


# ---- Part 1: drop-in boundary -------------------------------------------------------------------
Pixel_HSV = _struct("Pixel_HSV", [("parent_id", C.c_int), ("h", C.c_double), ("s", C.c_double), ("v", C.c_double)])
# the reference declares height/width as c_uint; same size and layout as the C `int`
Image_RGB = _struct("Image_RGB", [("height", C.c_uint), ("width", C.c_uint), ("r", _dp), ("g", _dp), ("b", _dp)])
Image_PGM = _struct("Image_PGM", [("height", C.c_uint), ("width", C.c_uint), ("data", _dp)])
RGB_Statistics = _struct("RGB_Statistics", [(k, C.c_double) for k in ("Br", "Bg", "Bb", "Cr", "Cg", "Cb")])
Crop_Boundaries = _struct("Crop_Boundaries", [("N", C.c_int), ("top", _ip), ("bottom", _ip), ("left", _ip), ("right", _ip)])
Color_Palette = _struct("Color_Palette", [("N", C.c_int), ("averages", C.POINTER(Pixel_HSV)), ("percentages", _dp)])
Blur_Vector = _struct("Blur_Vector", [("angle", C.c_int), ("magnitude", C.c_float)])
Blur_Vector_Group = _struct("Blur_Vector_Group", [("len_vectors", C.c_int), ("blur_vectors", C.POINTER(Blur_Vector))])
Sharpnesses = _struct("Sharpnesses", [("N", C.c_int), ("sharpness", _dp)])


def _bin_values(self):
    """bins[angle][radius] as nested lists (same helper as the reference's Blur_Profile.get_bin_values)."""
    return [[self.bins[a][r] for r in range(self.num_radius_bins)] for a in range(self.num_angle_bins)]


Blur_Profile = _struct("Blur_Profile",
                       [("num_angle_bins", C.c_int), ("num_radius_bins", C.c_int), ("angle_bin_size", C.c_int),
                        ("radius_bin_size", C.c_int), ("bins", C.POINTER(_dp))],
                       {"get_bin_values": _bin_values})
Full_Report_Data = _struct("Full_Report_Data",
                           [("rgb_stats", C.POINTER(RGB_Statistics)), ("color_palette", C.POINTER(Color_Palette)),
                            ("blur_profile", C.POINTER(Blur_Profile)), ("blur_vectors", C.POINTER(Blur_Vector_Group)),
                            ("average_saturation", C.c_double), ("sharpness", C.POINTER(Sharpnesses))])

# ---- Part 2: batch interface ---------------------------------------------------------------------
phd_params = _struct("phd_params",
                     [("h_partitions", C.c_int), ("s_partitions", C.c_int), ("v_partitions", C.c_int),
                      ("black_thresh", C.c_double), ("gray_thresh", C.c_double), ("coverage_thresh", C.c_double),
                      ("linked_list_size", C.c_int), ("downsample_rate", C.c_int), ("radius_partitions", C.c_int),
                      ("angle_partitions", C.c_int), ("quantity_weight", C.c_float),
                      ("saturation_value_weight", C.c_float), ("fft_streak_thresh", C.c_double),
                      ("magnitude_thresh", C.c_double), ("blur_cutoff_ratio_denom", C.c_int)])
phd_flat_head = _struct("phd_flat_head",
                        [("rgb_stats", C.c_double * 6), ("average_saturation", C.c_double), ("max_power", C.c_double),
                         ("dropped_pixels", C.c_longlong), ("palette_n", C.c_int), ("tie_groups", C.c_int),
                         ("n_sharpness", C.c_int), ("angle_bin_size", C.c_int), ("radius_bin_size", C.c_int),
                         ("num_angle_bins", C.c_int), ("num_radius_bins", C.c_int), ("status", C.c_int),
                         ("blur_vec_angle", C.c_int * 10), ("blur_vec_mag", C.c_float * 10)])
phd_flat_layout = _struct("phd_flat_layout",
                          [("record_bytes", C.c_size_t), ("off_palette_hsv", C.c_size_t),
                           ("off_palette_pct", C.c_size_t), ("off_parent_ids", C.c_size_t),
                           ("off_blur_bins", C.c_size_t), ("off_sharpness", C.c_size_t), ("T", C.c_int),
                           ("na", C.c_int), ("nr", C.c_int), ("max_boxes", C.c_int)])

# sizes the C side must agree with (checked by tests/test_abi.py)
ABI_SIZES = {"Image_RGB": 32, "Image_PGM": 16, "Pixel_HSV": 32, "Crop_Boundaries": 40, "RGB_Statistics": 48,
             "Color_Palette": 24, "Blur_Profile": 24, "Blur_Vector": 8, "Blur_Vector_Group": 16, "Sharpnesses": 16,
             "Full_Report_Data": 48}
