"""Comparison rules shared by the parity tests.

Tolerances (BASELINE.json north_star: bit-exact integers, <= 1e-4 relative for floating point).  What the
implementation actually achieves is far tighter, and the tests pin that so regressions show:

  integer outputs (palette N, parent order, angle/radius bin sizes, blur-vector angles, group counts,
  bin-id map, tie/dropped pixel counts)                        : exact
  palette percentages (count / P)                              : 1e-15 absolute (same integers, same division)
  rgb_stats, sharpness, palette v (exact integer sums)         : 1e-9 relative
  average_saturation, palette s (per-pixel s rounded to 2^-20) : 2e-6 relative
  palette hue (degrees, circular; per-pixel hue rounded to
      2^-20 of a half hue bin, i.e. 1e-5 degree at h=18)       : 1e-5 absolute
  blur-profile bins (FP32 transform vs the reference's FP64)   : 1e-4 relative with a 2e-6 absolute floor
  blur-vector magnitudes (k / nr as float)                     : exact

FFT magnitudes.  The transform runs in FP32 (SURVEY.md section 8d sanctions an FP32 complex intermediate); its error
is ABSOLUTE in nature -- about 1e-7 of the spectrum's rms level per coefficient -- so a coefficient far below the rms
level cannot be relative-accurate to 1e-4 in any FP32 transform.  tests/test_gpu_parity.py therefore asserts
    |P - P_ref| <= 1e-4 * (P_ref + mean(P_ref))                  (P = |X|^2; the per-coefficient bound relative to the
                                                                 coefficient PLUS the mean power level)
and, separately, the north star's plain per-coefficient figure on the coefficients that carry the result:
    |P - P_ref| / P_ref <= 1e-4 for every coefficient with P_ref >= 1e-3 * mean(P_ref)
and it REPORTS (fft_error_report, printed with -s / -rP and kept in the test's user properties) the distribution of the
plain per-coefficient relative error of the MAGNITUDE |X| (median / p99 / max over all coefficients) and the number of
coefficients on the other side of the `p < 1 -> 0` threshold of pgm_normalize_fft (src/fft_processing.c:188-193) than in
the float64 transform.  What reaches the report -- the blur-profile bins, means of ln p over hundreds of coefficients --
is held to 1e-4 relative by assert_report_close.
"""
import numpy as np

RTOL_STATS = 1e-9
RTOL_SAT = 2e-6
ATOL_HUE = 1e-5
RTOL_BINS = 1e-4
ATOL_BINS = 2e-6


def rel_err(a, b, floor=1e-300):
    """Relative error; positions where BOTH sides are NaN (e.g. 0/0 of an empty parent, which the reference
    also produces) count as equal, a NaN on one side only as infinitely wrong."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    e = np.where(both, 0.0, e)
    return np.where(np.isnan(e), np.inf, e)


def planes_for(kind, seed, W, H):
    """Planes of doubles that are NOT k/255 (the general-input route): 'test_rgb' is the reference's own
    create_test_rgb (src/debug.c:53-60), 'random' numpy's PCG64 stream (platform independent), 'k65535' 16-bit values."""
    n = W * H
    if kind == "test_rgb":
        i = np.arange(n, dtype=np.float64)
        return tuple((1.0 / (i * m + 1.0)).reshape(H, W) for m in (1.0, 2.0, 4.0))
    rng = np.random.default_rng(seed)
    if kind == "random":
        return tuple(rng.random((H, W)) for _ in range(3))
    if kind == "k65535":
        return tuple(rng.integers(0, 65536, (H, W)).astype(np.float64) / 65535.0 for _ in range(3))
    raise ValueError(kind)



def fft_error_report(pw, ref):
    """pw: FP32 power spectrum of the product path, ref: float64 power spectrum (same shape).  Returns a dict with the
    per-coefficient relative error of the magnitude |X| (median, p99, max, and max over the coefficients above 1e-3 of
    the mean power) and the count of `p < 1` threshold flips."""
    pw = np.asarray(pw, np.float64)
    ref = np.asarray(ref, np.float64)
    mag, mref = np.sqrt(pw), np.sqrt(ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(mag - mref) / mref
    rel = np.where(mref == 0, np.where(mag == 0, 0.0, np.inf), rel)
    big = ref >= 1e-3 * ref.mean()
    flips = int(np.count_nonzero((pw < 1.0) != (ref < 1.0)))
    return dict(n=int(ref.size), median=float(np.median(rel)), p99=float(np.quantile(rel, 0.99)), max=float(rel.max()),
                max_significant=float(rel[big].max()) if big.any() else 0.0,
                n_significant=int(big.sum()), threshold_flips=flips,
                max_power_rel=float((np.abs(pw - ref) / (ref + ref.mean())).max()))


def assert_report_close(got, want, what=""):
    """got / want: oracle.binding.Report-like objects (numpy fields)."""
    assert np.all(rel_err(got.rgb_stats, want.rgb_stats, 1e-12) < RTOL_STATS), f"{what}: rgb_stats"
    assert rel_err(got.average_saturation, want.average_saturation, 1e-12) < RTOL_SAT, f"{what}: average_saturation"
    assert len(got.palette_pct) == len(want.palette_pct), f"{what}: palette N {len(got.palette_pct)} != {len(want.palette_pct)}"
    assert np.max(np.abs(got.palette_pct - want.palette_pct), initial=0) <= 1e-15, f"{what}: palette percentages"
    if len(want.palette_pct):
        dh = rel_err(got.palette_hsv[:, 0], want.palette_hsv[:, 0], 1.0) * np.maximum(np.abs(np.nan_to_num(want.palette_hsv[:, 0])), 1.0)
        dh = np.minimum(dh, np.abs(360 - dh))
        assert np.max(dh) < ATOL_HUE, f"{what}: palette hue {np.max(dh)}"
        assert np.all(rel_err(got.palette_hsv[:, 1], want.palette_hsv[:, 1], 1e-12) < RTOL_SAT), f"{what}: palette s"
        assert np.all(rel_err(got.palette_hsv[:, 2], want.palette_hsv[:, 2], 1e-12) < RTOL_STATS), f"{what}: palette v"
    assert got.angle_bin_size == want.angle_bin_size and got.radius_bin_size == want.radius_bin_size, f"{what}: bin sizes"
    d = np.abs(got.blur_bins - want.blur_bins)
    assert np.all(d <= ATOL_BINS + RTOL_BINS * np.abs(want.blur_bins)), f"{what}: blur bins max abs {d.max()}"
    assert np.array_equal((got.blur_bins == 0), (want.blur_bins == 0)), f"{what}: empty-bin pattern"
    assert np.array_equal(got.blur_vec_angle, want.blur_vec_angle), f"{what}: blur vector angles"
    assert np.array_equal(got.blur_vec_mag, want.blur_vec_mag), f"{what}: blur vector magnitudes"
    if want.sharpness is None:
        assert got.sharpness is None or len(got.sharpness) == 0, f"{what}: unexpected sharpness"
    else:
        assert got.sharpness is not None and len(got.sharpness) == len(want.sharpness), f"{what}: sharpness count"
        assert np.all(rel_err(got.sharpness, want.sharpness, 1e-12) < RTOL_STATS), f"{what}: sharpness"


def report_from_batch(b, i):
    """One record of a photohive_dsp_b200.batch.BatchReports as an oracle.binding.Report."""
    from oracle.binding import Report
    n = int(b.palette_n[i])
    return Report(rgb_stats=b.rgb_stats[i].copy(), average_saturation=float(b.average_saturation[i]),
                  palette_hsv=b.palette_hsv[i, :n].copy(), palette_pct=b.palette_pct[i, :n].copy(),
                  blur_bins=b.blur_bins[i].copy(), angle_bin_size=int(b.angle_bin_size[i]),
                  radius_bin_size=int(b.radius_bin_size[i]), blur_vec_angle=b.blur_vec_angle[i].copy(),
                  blur_vec_mag=b.blur_vec_mag[i].copy(),
                  sharpness=None if b.sharpness is None else b.sharpness[i].copy(),
                  extra=dict(parent_ids=b.parent_ids[i, :n].copy(), tie_groups=int(b.tie_groups[i]),
                             dropped_pixels=int(b.dropped_pixels[i])))


def golden_report(golden, name):
    from oracle.binding import Report
    ints = golden.field(name, "ints")
    return Report(rgb_stats=golden.field(name, "rgb_stats"), average_saturation=float(golden.field(name, "average_saturation")),
                  palette_hsv=golden.field(name, "palette_hsv"), palette_pct=golden.field(name, "palette_pct"),
                  blur_bins=golden.field(name, "blur_bins"), angle_bin_size=int(ints[0]), radius_bin_size=int(ints[1]),
                  blur_vec_angle=golden.field(name, "blur_vec_angle"), blur_vec_mag=golden.field(name, "blur_vec_mag"),
                  sharpness=golden.field(name, "sharpness"))


def boxes_array(meta_boxes, n=1):
    if not meta_boxes:
        return None
    one = [[b["top"], b["bottom"], b["left"], b["right"]] for b in meta_boxes]
    return np.array([one] * n, np.int32)
