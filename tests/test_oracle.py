"""Pins the CPU oracle (no GPU needed).

1. against the committed golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py);
2. against the reference itself, live, wherever its compiled form is available (oracle/_ref);
3. its FFT against numpy, its generators against an independent restatement of SURVEY.md B.2;
4. the reference's own three rejection tests (src/test/test.c:87-135).
"""
import numpy as np
import pytest

from oracle.binding import make_params
from parity import assert_report_close, golden_report

SMALL = 700 * 500


def _oracle_report(oracle, golden, name, nthreads=4):
    m = golden.meta[name]
    img = golden.image(oracle, name)
    return oracle.report(img, make_params(**m["params"]), boxes=m["boxes"], nthreads=nthreads)


@pytest.mark.parametrize("name", ["g1_small", "g0_small", "g2_small", "g1_odd", "g1_odd2", "g0_fine", "g1_down5", "g1_list50",
                                  "g2_cov1", "g1_1080p", "g0_1080p", "g2_1080p"])
def test_oracle_matches_reference_golden(oracle, golden, name):
    got = _oracle_report(oracle, golden, name)
    want = golden_report(golden, name)
    assert_report_close(got, want, name)
    # the oracle is double precision end to end: bins must agree far below the GPU tolerance
    assert np.max(np.abs(got.blur_bins - want.blur_bins)) < 1e-11


@pytest.mark.parametrize("name", ["f64_test_rgb", "f64_random", "f64_random_ties", "f64_k65535"])
def test_oracle_matches_reference_golden_on_general_doubles(oracle, golden, name):
    """Planes of doubles that are not k/255 (create_test_rgb of src/debug.c:53, random doubles, 16-bit values): the
    oracle's double-precision path against the fixtures the unmodified reference wrote."""
    m = golden.meta[name]
    got = oracle.report(None, make_params(**m["params"]), boxes=m["boxes"], nthreads=4, planes=golden.planes(name))
    want = golden_report(golden, name)
    assert_report_close(got, want, name)
    assert np.max(np.abs(got.blur_bins - want.blur_bins)) < 1e-11


def test_survey_known_answers(golden):
    """SURVEY.md B.3 values (reference at -O0, defaults, 1920x1080) are what the fixtures hold."""
    g1 = golden_report(golden, "g1_1080p")
    assert np.allclose(g1.rgb_stats, [0.50154320231181415, 0.50175441025669254, 0.52502899948873205,
                                      0.288850021322691, 0.28884297165577916, 0.212869547900532], rtol=1e-13)
    assert abs(g1.average_saturation - 0.50058217729694376) < 1e-14
    assert len(g1.palette_pct) == 16 and abs(g1.palette_pct.sum() - 0.993465470679) < 1e-11
    assert np.allclose(g1.palette_hsv[0], [321.697050115, 0.760302527759, 0.860633002134], rtol=1e-9)
    assert np.allclose(g1.blur_bins[0, :4], [0.6543027841, 0.4311515922, 0.3977491733, 0.3925834322], rtol=1e-9)
    assert (g1.angle_bin_size, g1.radius_bin_size) == (2, 27)
    g0 = golden_report(golden, "g0_1080p")
    assert len(g0.palette_pct) == 79 and abs(g0.palette_pct.sum() - 0.995223765432) < 1e-11
    g2 = golden_report(golden, "g2_1080p")
    assert len(g2.palette_pct) == 18 and abs(g2.palette_pct.sum() - 0.963940489969) < 1e-11
    assert g2.blur_vec_angle[0] == 2 and abs(g2.blur_vec_mag[0] - 0.15) < 1e-7 and not g2.blur_vec_angle[1:].any()


@pytest.mark.parametrize("case", [
    dict(kind=1, seed=21, W=500, H=380, kw={}),
    dict(kind=0, seed=22, W=384, H=512, kw=dict(h_partitions=12, s_partitions=3, v_partitions=2, black_thresh=0.2)),
    dict(kind=2, seed=23, W=450, H=350, kw=dict(coverage_thresh=0.8, linked_list_size=7)),
    dict(kind=1, seed=24, W=700, H=420, kw=dict(downsample_rate=2, quantity_weight=0.5, saturation_value_weight=0.5)),
    dict(kind=0, seed=25, W=360, H=360, kw=dict(radius_partitions=10, angle_partitions=18, blur_cutoff_ratio_denom=3)),
])
def test_oracle_matches_reference_live(oracle, reference, case, capfd):
    img = oracle.generate(case["kind"], case["seed"], case["W"], case["H"])
    W, H = case["W"], case["H"]
    boxes = [dict(top=3, bottom=H // 2, left=5, right=W // 3), dict(top=H // 4, bottom=H, left=W // 2, right=W)]
    p = make_params(**case["kw"])
    want = reference.report(img, p, boxes=boxes)
    got = oracle.report(img, p, boxes=boxes, nthreads=2)
    capfd.readouterr()  # the reference prints its stage timings
    assert_report_close(got, want, str(case))
    assert np.max(np.abs(got.blur_bins - want.blur_bins)) < 1e-11


def test_reference_rejection_rules(oracle):
    # src/test/test.c:87-135 and src/utilities.c:64-87
    assert oracle.rejects(120000, 10000)          # too many pixels
    assert oracle.rejects(2001, 400) and oracle.rejects(400, 2001)   # aspect ratio beyond 5:1
    assert oracle.rejects(349, 350)               # a side below 350
    assert not oracle.rejects(350, 350) and not oracle.rejects(1920, 1080) and not oracle.rejects(12000, 10000)


def test_reference_rejects_live(reference, capfd):
    from oracle.binding import make_params as mp
    for (W, H) in [(2001, 400), (400, 2001), (349, 350)]:
        planes = tuple(np.zeros((H, W)) for _ in range(3))
        assert reference.report(None, mp(), planes=planes) is None
    capfd.readouterr()


def test_oracle_fft_against_numpy(oracle):
    """Power spectrum of the oracle (mixed radix 357 = 3*7*17, Bluestein 401) vs numpy's rfft2."""
    img = oracle.generate(1, 77, 401, 357)
    rep = oracle.report(img, stages=4, want_intermediates=True)
    gray = rep.extra["gray"].reshape(357, 401)
    avg = rep.rgb_stats[:3].sum() / 3.0
    want = np.abs(np.fft.rfft2(gray - avg)) ** 2
    got = rep.extra["power"].reshape(357, 201)
    assert np.max(np.abs(got - want)) / np.max(want) < 1e-12
    assert np.median(np.abs(got - want) / np.maximum(want, 1e-30)) < 1e-10


def _lcg_stream(seed, n):
    out = np.empty(n, np.uint32)
    s = seed
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out[i] = s >> 33
    return out


def test_generators_follow_survey_b2(oracle):
    W, H = 37, 23
    g0 = oracle.generate(0, 12345, W, H)
    assert np.array_equal(g0.ravel(), (_lcg_stream(12345, W * H * 3) & 255).astype(np.uint8))
    g1 = oracle.generate(1, 12345, W, H)
    n = (_lcg_stream(12345, W * H) & 15).astype(np.int64).reshape(H, W)
    x, y = np.meshgrid(np.arange(W), np.arange(H))
    assert np.array_equal(g1[:, :, 0], ((x * 255 // W + n) & 255).astype(np.uint8))
    assert np.array_equal(g1[:, :, 1], ((y * 255 // H + n) & 255).astype(np.uint8))
    assert np.array_equal(g1[:, :, 2], (((x + y) * 255 // (W + H) + n) & 255).astype(np.uint8))


def test_tie_path_drops_pixels(oracle, golden):
    """percentages sum below 1 when the tail node overflows; never above 1."""
    for name in ["g1_small", "g1_list50", "g2_small"]:
        rep = _oracle_report(oracle, golden, name)
        m = golden.meta[name]
        P = m["W"] * m["H"]
        assert abs(rep.palette_pct.sum() - (1 - rep.extra["dropped_pixels"] / P)) < 1e-12
        assert rep.extra["dropped_pixels"] > 0 and rep.extra["tie_groups"] > 0
