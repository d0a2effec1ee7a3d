"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle, the committed
reference golden vectors, and size-independent properties at BASELINE.json's full sizes.  Run with -m gpu."""
import ctypes as C
import zlib

import numpy as np
import pytest

from oracle.binding import make_params as omake
from parity import assert_report_close, boxes_array, fft_error_report, golden_report, rel_err, report_from_batch
from photohive_dsp_b200.batch import flat_layout, make_params

pytestmark = pytest.mark.gpu

FINE = dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99)


# ---- discrete stages: bit exact --------------------------------------------------------------------
@pytest.mark.parametrize("kw", [{}, FINE, dict(h_partitions=12, s_partitions=3, v_partitions=5, black_thresh=0.15, gray_thresh=0.2),
                                dict(h_partitions=24, s_partitions=1, v_partitions=1)])
def test_group_id_of_all_2pow24_colours(ctx, oracle, kw):
    """the palette group of every 24-bit colour equals the reference arithmetic, bit for bit --
    both for the product path (integer fast path + FP64 edge path) and for the plain FP64 transcription."""
    want = oracle.group_sweep(omake(**kw))
    got = ctx.debug_group_sweep(make_params(**kw))
    assert np.array_equal(got, want), f"fast path: {np.count_nonzero(got != want)} colours land in another group"
    got = ctx.debug_group_sweep(make_params(**kw), exact=True)
    assert np.array_equal(got, want), f"FP64 path: {np.count_nonzero(got != want)} colours land in another group"


@pytest.mark.parametrize("shape", [(1920, 1080), (3840, 2160), (6000, 4000), (405, 357), (1080, 1920), (350, 350), (752, 502)])
def test_polar_bin_map_is_identical(ctx, oracle, shape):
    """bin ids (truncated PI, Newton sqrt, bottom-half row rule) and bin populations."""
    m, c = ctx.debug_bin_map(*shape)
    mo, co = oracle.bin_map(*shape)
    assert np.array_equal(m, mo) and np.array_equal(c, co)


@pytest.mark.parametrize("kind,kw", [(0, {}), (1, {}), (1, dict(downsample_rate=3)), (0, FINE)])
def test_group_counts_bit_exact(ctx, oracle, kind, kw):
    img = oracle.generate(kind, 4242 + kind, 1280, 720)
    got = ctx.debug_group_counts(img, make_params(**kw))
    want = oracle.report(img, omake(**kw), stages=1).extra["group_counts"]
    assert np.array_equal(got, want)
    ds = kw.get("downsample_rate", 1)
    assert got.sum() == (1280 // ds) * (720 // ds)


# ---- full reports ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["g1_small", "g0_small", "g2_small", "g1_odd", "g1_odd2", "g0_fine", "g1_down5", "g1_list50", "g2_cov1",
                                  "g1_1080p", "g0_1080p", "g2_1080p"])
def test_report_matches_reference_golden(ctx, oracle, golden, name):
    """Against outputs of the UNMODIFIED reference (tests/golden/make_golden.py)."""
    m = golden.meta[name]
    img = golden.image(oracle, name)
    b = ctx.get_reports(img[None], boxes=boxes_array(m["boxes"]), params=make_params(**m["params"]))
    assert_report_close(report_from_batch(b, 0), golden_report(golden, name), name)


@pytest.mark.parametrize("W,H,kind,kw", [
    (1920, 1080, 1, {}), (1000, 750, 0, {}), (1280, 960, 2, dict(linked_list_size=16)),
    (700, 525, 1, dict(h_partitions=9, s_partitions=3, v_partitions=4, coverage_thresh=0.9)),
    (1024, 768, 0, dict(downsample_rate=2, radius_partitions=16, angle_partitions=36)),
    (3840, 2160, 1, {}),   # BASELINE config 2: 4K with four salient boxes
    (1080, 1920, 0, {}),   # portrait 1080p: the row kernel's 8-pixel staging segments
    (1600, 900, 1, {}),    # 900-point columns: the specialised column kernel with a rounded-up last bin-id slice
    (752, 502, 1, {}),     # 2^4*47 x 2*251: prime factors served by the O(p^2) butterfly of the generic FFT kernels
    (1008, 572, 0, {}),    # 2^4*3^2*7 x 2^2*11*13: register butterflies for the primes 7, 11 and 13
    (646, 456, 1, {}),     # 2*17*19 x 2^3*3*19
    (1031, 523, 2, {}),    # both sides prime (1031 > the 1021 that earlier builds refused)
    # saliencies closer than 1 apart: the truncating comparator calls them equal (insertion-sort replay)
    (800, 600, 0, dict(quantity_weight=0.0, saturation_value_weight=1e-5)),
    (800, 600, 1, dict(quantity_weight=1e-6, saturation_value_weight=1e-6, coverage_thresh=0.5)),
    # saliencies a fraction of a pixel count apart: many short chains of "equal" neighbours (chain repair of
    # k_palette_select), also with the fine palette's 871 groups
    (800, 600, 0, dict(quantity_weight=3.7e-4, saturation_value_weight=2.3e-4)),
    (800, 600, 1, dict(quantity_weight=1.3e-3, saturation_value_weight=4.1e-4, coverage_thresh=0.9)),
    (800, 600, 0, dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99, quantity_weight=2.9e-3,
                       saturation_value_weight=1.7e-3)),
    # saliencies beyond 2^31: the float->int conversion overflows to INT_MIN
    (800, 600, 0, dict(quantity_weight=30000.0, saturation_value_weight=50000.0)),
    (640, 480, 2, dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99, linked_list_size=3)),
])
def test_report_matches_oracle(ctx, oracle, W, H, kind, kw):
    img = oracle.generate(kind, 1000 + W + kind, W, H)
    boxes = [dict(top=H * i // 8, bottom=H * i // 8 + H // 4, left=W * i // 8, right=W * i // 8 + W // 4) for i in range(4)]
    want = oracle.report(img, omake(**kw), boxes=boxes, nthreads=8)
    b = ctx.get_reports(img[None], boxes=boxes_array(boxes), params=make_params(**kw))
    got = report_from_batch(b, 0)
    assert_report_close(got, want, f"{W}x{H} kind {kind} {kw}")
    assert np.array_equal(got.extra["parent_ids"], want.extra["parent_ids"])
    assert got.extra["tie_groups"] == want.extra["tie_groups"]
    assert got.extra["dropped_pixels"] == want.extra["dropped_pixels"]


@pytest.mark.parametrize("W,H,kind,kw,nbox", [
    (6000, 4000, 1, {}, 0),       # BASELINE config 4: 24 MP; saliencies pass 2^31 on their own
    (6000, 4000, 0, {}, 2),       # 24 MP noise: 79 parents, 5,860 chunks per image, boxes on the large image
    (1920, 1080, 0, FINE, 0),     # BASELINE config 5: h36 s4 v6 cov .99 at 1080p (the 512-thread front-end variant)
    (1920, 1080, 1, FINE, 4),
    (1920, 1080, 2, FINE, 0),
    (4000, 6000, 1, dict(downsample_rate=2), 0),  # portrait 24 MP, downsampled HSV grid
    (7680, 4320, 1, {}, 0),       # 33 MP: two groups' saliencies exceed 2^31 without any help from the weights
    (13000, 2600, 1, {}, 1),      # a side beyond the shared-memory row FFT: four-step transform through HBM
    (2602, 13003, 0, {}, 0),      # a prime column length beyond the shared-memory column FFT (Bluestein through HBM)
])
def test_full_report_at_baseline_sizes(ctx, oracle, W, H, kind, kw, nbox):
    """BASELINE configs 4 and 5 as FULL reports against the oracle (the stages whose indexing changes with
    the image size -- chunk counts, work lists, natural saliency overflow, the 6000/4000-point FFT epilogues -- and the
    fine-palette front end at 1080p were only covered through invariants / small images)."""
    img = oracle.generate(kind, 2024 + kind, W, H)
    boxes = [dict(top=H * i // 8, bottom=H * i // 8 + H // 4, left=W * i // 8, right=W * i // 8 + W // 4)
             for i in range(nbox)] or None
    want = oracle.report(img, omake(**kw), boxes=boxes, nthreads=16)
    b = ctx.get_reports(img[None], boxes=boxes_array(boxes), params=make_params(**kw))
    got = report_from_batch(b, 0)
    assert_report_close(got, want, f"{W}x{H} kind {kind} {kw}")
    assert np.array_equal(got.extra["parent_ids"], want.extra["parent_ids"])
    assert got.extra["tie_groups"] == want.extra["tie_groups"]
    assert got.extra["dropped_pixels"] == want.extra["dropped_pixels"]
    counts = ctx.debug_group_counts(img, make_params(**kw))
    assert np.array_equal(counts, oracle.report(img, omake(**kw), stages=1).extra["group_counts"])
    if (W, H) == (7680, 4320):
        # the comparator's float -> int overflow (cvttss2si -> INT_MIN) is live here with the default weights
        p = omake(**kw)
        g = np.arange(p.h_partitions * p.s_partitions * p.v_partitions)
        sc = ((g // p.v_partitions) % p.s_partitions + 0.5) * (1 - p.gray_thresh) / p.s_partitions + p.gray_thresh
        vc = (g % p.v_partitions + 0.5) * (1 - p.black_thresh) / p.v_partitions + p.black_thresh
        sal = counts[:len(g)] * (p.quantity_weight + p.saturation_value_weight * sc * vc) * 1000
        assert (sal > 2.0 ** 31).sum() >= 1


def _exceptional_colour_image(hp, W, H, seed, gray_fraction):
    """Pixels whose hue sits EXACTLY on a half-hue-bin boundary (where the reference's double rounding decides the bin
    and the side of the wrap seam), plus a gray area so that a gray / black parent (seam at 180 degrees) is selected."""
    c = np.arange(1 << 24, dtype=np.int64)
    R, G, B = c & 255, (c >> 8) & 255, c >> 16
    mx, mn = np.maximum(R, np.maximum(G, B)), np.minimum(R, np.minimum(G, B))
    q = mx - mn
    is_r = R == mx
    is_g = (~is_r) & (G == mx)
    p = np.where(is_r, G - B, np.where(is_g, B - R, R - G))
    num2 = 120 * (np.where(is_r, 0, np.where(is_g, 2, 4)) * q + p)
    num2 = np.where(num2 < 0, num2 + 720 * q, num2)
    exc = c[(q > 0) & (num2 % np.maximum((360 // hp) * q, 1) == 0)]
    rng = np.random.default_rng(seed)
    pick = exc[rng.integers(0, len(exc), W * H)]
    img = np.stack([pick & 255, (pick >> 8) & 255, pick >> 16], -1).astype(np.uint8).reshape(H, W, 3)
    n_gray = int(H * gray_fraction)
    img[:n_gray] = rng.integers(40, 200, (n_gray, W, 1)).astype(np.uint8)  # R == G == B: saturation 0
    return img


@pytest.mark.parametrize("hp,cov,frac", [(18, 0.5, 0.3), (18, 0.95, 0.0), (9, 0.4, 0.3), (36, 0.3, 0.5), (12, 0.7, 0.1)])
def test_boundary_colours_follow_the_reference_rounding(ctx, oracle, hp, cov, frac):
    """Every coloured pixel is an 'exceptional' one (pixel_cells.cuh): exact on a half-bin boundary.  Few parents are
    selected, so far-away groups join them and their pixels cross the wrap seams of calculate_avg_hsv."""
    kw = dict(h_partitions=hp, coverage_thresh=cov)
    img = _exceptional_colour_image(hp, 640, 480, 77 + hp, frac)
    want = oracle.report(img, omake(**kw), nthreads=8)
    got = report_from_batch(ctx.get_reports(img[None], params=make_params(**kw)), 0)
    assert_report_close(got, want, f"exceptional colours hp={hp}")
    assert np.array_equal(got.extra["parent_ids"], want.extra["parent_ids"])
    assert got.extra["dropped_pixels"] == want.extra["dropped_pixels"]
    counts = ctx.debug_group_counts(img, make_params(**kw))
    assert np.array_equal(counts, oracle.report(img, omake(**kw), stages=1).extra["group_counts"])


def test_batch_is_deterministic_and_position_independent(ctx, oracle):
    """Integer accumulators: a record does not depend on where the image sits in a batch, nor on the run."""
    imgs = np.stack([oracle.generate(k % 3, 50 + k, 640, 480) for k in range(7)])
    a = ctx.get_reports(imgs)
    b = ctx.get_reports(imgs[::-1].copy())
    c = ctx.get_reports(imgs)
    assert np.array_equal(a.raw, c.raw)
    assert np.array_equal(a.raw, b.raw[::-1])


def test_device_resident_input_equals_host_input(ctx, oracle):
    import torch
    imgs = np.stack([oracle.generate(k, 9 + k, 800, 600) for k in range(3)])
    host = ctx.get_reports(imgs)
    dev = ctx.get_reports(torch.from_numpy(imgs).cuda())
    assert np.array_equal(host.raw, dev.raw)


@pytest.mark.parametrize("env", [{"PHD_NO_TMA_STORE": "1"}, {"PHD_ROWS_KTMA": "256"}, {"PHD_ROWS_KTMA": "100000"}])
@pytest.mark.parametrize("W,H", [(1920, 1080), (3840, 2160), (640, 480)])
def test_row_output_paths_give_identical_records(ctx, oracle, env, W, H):
    """The row kernel's transposed spectrum leaves through a TMA tensor store, through direct stores, or split between
    them (fft.cu, rows_walk); which path takes which spectrum column must not change a single bit of a record.  The
    choice is read once per process, so the other paths run in a child process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, zlib, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "from oracle.binding import Oracle\n"
        "from photohive_dsp_b200.batch import Context\n"
        f"imgs = np.stack([Oracle().generate(k % 3, 77 + k, {W}, {H}) for k in range(3)])\n"
        "print(zlib.crc32(Context(0).get_reports(imgs).raw.tobytes()))\n")
    imgs = np.stack([oracle.generate(k % 3, 77 + k, W, H) for k in range(3)])
    want = zlib.crc32(ctx.get_reports(imgs).raw.tobytes())
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert int(out.stdout.strip().splitlines()[-1]) == want, env


def test_two_devices_in_one_process(ctx, oracle):
    """One context per GPU inside ONE process (the library allows it): same records from both devices."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from photohive_dsp_b200.batch import Context
    imgs = np.stack([oracle.generate(k % 3, 500 + k, 1920, 1080) for k in range(2)] )
    a = ctx.get_reports(imgs)
    other = Context(1)
    try:
        b = other.get_reports(imgs)
        c = other.get_reports(torch.from_numpy(imgs).to("cuda:1"))
    finally:
        other.close()
    assert np.array_equal(a.raw, b.raw) and np.array_equal(a.raw, c.raw)


def test_multi_gpu_entry_point_of_the_c_abi(ctx, oracle):
    """phd_get_reports_u8_multi: contiguous ranges of one host batch on every visible GPU, records gathered by the join.
    On a one-GPU box the call still runs (one range); with two or more the ranges land on different devices."""
    import torch
    from photohive_dsp_b200.batch import MultiContext
    ndev = torch.cuda.device_count()
    imgs = np.stack([oracle.generate(k % 3, 800 + k, 800, 600) for k in range(7)])   # 7: uneven ranges
    boxes = np.array([[[0, 300, 0, 400], [100, 600, 300, 800]]] * 7, np.int32)
    want = ctx.get_reports(imgs, boxes=boxes)
    for devices in ([0], list(range(ndev)), list(range(ndev))[::-1]):
        mc = MultiContext(devices)
        try:
            got = mc.get_reports(imgs, boxes=boxes)
            pinned = torch.from_numpy(imgs).pin_memory()
            lay = got.layout
            rec = np.empty((7, lay.record_bytes), np.uint8)
            mc.get_reports_raw(pinned.data_ptr(), 7, 800, 600, 800 * 600 * 3, make_params(), rec.ctypes.data,
                               boxes.ctypes.data, 2)
        finally:
            mc.close()
        assert np.array_equal(got.raw, want.raw) and np.array_equal(rec, want.raw)
    # device buffers and repeated contexts are refused
    from photohive_dsp_b200.batch import PhotoHiveError
    mc = MultiContext([0])
    try:
        with pytest.raises(PhotoHiveError):
            mc.get_reports_raw(torch.from_numpy(imgs).cuda().data_ptr(), 7, 800, 600, 800 * 600 * 3, make_params(),
                               rec.ctypes.data)
    finally:
        mc.close()


def test_batch_entry_point_validates_boxes(ctx, oracle):
    """Boxes outside the image or inverted are refused (the reference's crop_pgm returns NULL and is then dereferenced);
    an empty box gives NaN like the reference, not an error."""
    from photohive_dsp_b200.batch import PhotoHiveError
    img = oracle.generate(1, 3, 640, 480)[None]
    for bad in ([0, 481, 0, 10], [-1, 10, 0, 10], [0, 10, 0, 641], [20, 10, 0, 10], [0, 10, 30, 20]):
        with pytest.raises(PhotoHiveError) as e:
            ctx.get_reports(img, boxes=np.array([[bad]], np.int32))
        assert e.value.code == 2
    b = ctx.get_reports(img, boxes=np.array([[[5, 5, 0, 10], [0, 480, 0, 640]]], np.int32))
    assert np.isnan(b.sharpness[0, 0]) and np.isfinite(b.sharpness[0, 1])


def test_pil_modes_follow_the_reference_conversion(oracle):
    """get_report takes the first three channels of np.array(pil_image) like the reference (utils.py:30-33): RGBA drops
    alpha, CMYK passes C, M, Y through, single-channel modes raise IndexError -- no silent .convert('RGB')."""
    from PIL import Image
    import photohive_dsp_b200 as P
    arr = oracle.generate(2, 21, 640, 480)
    rgb = P.get_report(Image.fromarray(arr, "RGB")).to_json()
    rgba = np.concatenate([arr, np.full((480, 640, 1), 77, np.uint8)], -1)
    assert P.get_report(Image.fromarray(rgba, "RGBA")).to_json() == rgb
    assert P.get_report(Image.fromarray(rgba, "CMYK")).to_json() == rgb      # C, M, Y bytes taken as they are
    with pytest.raises(IndexError):
        P.get_report(Image.fromarray(arr[:, :, 0], "L"))


def test_edge_images(ctx, oracle):
    """Flat, black, white and two-colour images: empty spectrum, single group, max==1 clamps."""
    H, W = 400, 560
    cases = {
        "black": np.zeros((H, W, 3), np.uint8),
        "white": np.full((H, W, 3), 255, np.uint8),
        "gray": np.full((H, W, 3), 128, np.uint8),
        "red": np.tile(np.array([255, 0, 0], np.uint8), (H, W, 1)),
    }
    half = np.zeros((H, W, 3), np.uint8)
    half[:, W // 2:] = [10, 200, 30]
    cases["two"] = half
    for name, img in cases.items():
        want = oracle.report(img, omake(), nthreads=4)
        got = report_from_batch(ctx.get_reports(img[None]), 0)
        # contrast of a constant image is rounding noise in the reference (1e-17) and exactly 0 here
        got.rgb_stats[3:] = np.where(np.abs(want.rgb_stats[3:]) < 1e-12, want.rgb_stats[3:], got.rgb_stats[3:])
        assert_report_close(got, want, name)


def test_sharpness_boxes_at_awkward_places(ctx, oracle):
    """Crop boxes of every byte phase, touching every image edge, one pixel wide / high, and around the 126-column and
    96-row strip sizes of the kernel, against the reference (zero padding at the CROP edge, src/filtering.c:81-107)."""
    W, H = 701, 523  # odd width: rows start at every byte phase
    img = oracle.generate(2, 8, W, H)
    boxes = [dict(top=0, bottom=H, left=0, right=W), dict(top=0, bottom=1, left=0, right=W),
             dict(top=H - 1, bottom=H, left=W - 1, right=W), dict(top=5, bottom=200, left=W - 127, right=W),
             dict(top=H - 97, bottom=H, left=3, right=129), dict(top=17, bottom=113, left=1, right=127),
             dict(top=17, bottom=114, left=2, right=129), dict(top=100, bottom=103, left=333, right=335),
             dict(top=0, bottom=H, left=350, right=351), dict(top=250, bottom=H, left=0, right=253)]
    want = oracle.report(img, omake(), boxes=boxes, nthreads=4)
    got = report_from_batch(ctx.get_reports(img[None], boxes=boxes_array(boxes)), 0)
    assert len(got.sharpness) == len(boxes)
    for i, (g, w) in enumerate(zip(got.sharpness, want.sharpness)):
        assert (np.isnan(g) and np.isnan(w)) or abs(g - w) <= 1e-9 * abs(w), (i, boxes[i], g, w)


# ---- drop-in entry point ---------------------------------------------------------------------------
def test_drop_in_entry_point_matches_oracle(oracle, capfd):
    """get_full_report_data on planar doubles k/255.0, exactly as core.py:442-469 calls it."""
    from oracle import binding
    from photohive_dsp_b200 import lib as L
    img = oracle.generate(1, 31337, 960, 540)
    boxes = [dict(top=10, bottom=200, left=20, right=400), dict(top=100, bottom=540, left=480, right=960)]
    lib = binding.bind_entry_points(C.CDLL(L.lib_path))
    rp = binding.call_entry_point(lib, binding.planes_from_u8(img), 960, 540, omake(), boxes)
    assert rp
    got = binding.unpack_full_report(rp)
    lib.free_full_report(C.byref(rp))
    assert not rp
    assert got.extra["len_vectors"] == 10
    assert_report_close(got, oracle.report(img, omake(), boxes=boxes, nthreads=4), "drop-in")
    # no boxes -> sharpness pointer is NULL (src/filtering.c:152-154)
    rp = binding.call_entry_point(lib, binding.planes_from_u8(img), 960, 540, omake(), None)
    assert rp and not rp.contents.sharpness
    lib.free_full_report(C.byref(rp))


def test_python_get_report_surface(oracle):
    """get_report / set_bounding_boxes / Report keep the reference's Python surface (core.py:23-119,388-515)."""
    import json
    import photohive_dsp_b200 as P
    img = oracle.generate(2, 5, 640, 480)
    bb = P.set_bounding_boxes([dict(top=0, bottom=240, left=0, right=320)])
    rep = P.get_report(img, salient_characters=bb)
    want = oracle.report(img, omake(), boxes=[dict(top=0, bottom=240, left=0, right=320)], nthreads=4)
    assert rep is not None and rep.rgb_stats.height == 480 and rep.rgb_stats.width == 640
    assert abs(rep.rgb_stats.Br - want.rgb_stats[0]) < 1e-12 and abs(rep.average_saturation - want.average_saturation) < 1e-7
    assert len(rep.blur_vectors) == 10 and len(rep.blur_profile.bins) == 72 and len(rep.blur_profile.bins[0]) == 40
    assert len(rep.color_palette.colors) == len(want.palette_pct) and len(rep.sharpnesses) == 1
    js = json.loads(rep.to_json())
    assert js["Height"] == 480 and abs(js["Sharpness 1:"] - want.sharpness[0]) < 1e-9 * abs(want.sharpness[0])
    vis = rep.generate_blur_profile_image()
    assert vis.size == (320, 480)
    assert P.get_report(np.zeros((349, 350, 3), np.uint8)) is None


def test_get_report_byte_route_equals_c_entry_point(oracle, monkeypatch):
    """get_report sends 8-bit images to the batch entry point (no float64 planes); the report must equal the one the
    reference's C entry point returns for the same image, field for field."""
    import photohive_dsp_b200 as P
    from photohive_dsp_b200 import core
    img = oracle.generate(1, 77, 640, 480)
    boxes = [dict(top=10, bottom=250, left=5, right=325), dict(top=100, bottom=480, left=300, right=640)]
    for bb in (None, boxes):
        def run():
            return P.get_report(img, salient_characters=None if bb is None else P.set_bounding_boxes(bb))
        monkeypatch.setattr(core, "_VIA_DOUBLES", False)
        fast = run()
        monkeypatch.setattr(core, "_VIA_DOUBLES", True)
        slow = run()
        assert fast.to_json() == slow.to_json()
        assert fast.color_palette.colors == slow.color_palette.colors
        assert fast.blur_profile.bins == slow.blur_profile.bins
        assert fast.sharpnesses == slow.sharpnesses and len(fast.sharpnesses) == (0 if bb is None else 2)
    # cases the byte route hands to the C entry point: empty box list, boxes outside the image
    monkeypatch.setattr(core, "_VIA_DOUBLES", False)
    rep = P.get_report(img, salient_characters=P.set_bounding_boxes([]))
    assert rep is not None and rep.sharpnesses == []
    assert P.get_report(img, salient_characters=P.set_bounding_boxes([dict(top=0, bottom=481, left=0, right=10)])) is None


@pytest.mark.parametrize("W,H,n", [(800, 600, 96), (1280, 720, 48), (640, 480, 96), (1024, 768, 48), (1920, 1080, 24),
                                   (752, 502, 24)])
def test_records_do_not_change_from_run_to_run(ctx, oracle, W, H, n):
    """All accumulators are integers and every shared-memory hand-over is fenced, so a batch gives the same bytes every
    time.  (Caught a race between the row FFT's thread groups: tolerance tests alone did not see it.)"""
    torch = pytest.importorskip("torch")
    base = np.stack([oracle.generate(k % 3, 4000 + k, W, H) for k in range(4)])
    imgs = torch.from_numpy(np.concatenate([base] * (n // 4))).cuda()
    p = make_params()
    lay = flat_layout(p, 0)
    outs = []
    for _ in range(4):
        rec = np.empty((n, lay.record_bytes), np.uint8)
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.ctypes.data)
        outs.append(rec)
    for rec in outs[1:]:
        assert (rec == outs[0]).all()
    for i in range(4, n):  # copies of one image give one record wherever they sit in the batch
        assert (outs[0][i] == outs[0][i % 4]).all()


def test_concurrent_callers_of_the_drop_in_entry_points(oracle):
    """The reference is not reentrant (globals, FFTW planner); the replacement serialises callers per context instead of
    corrupting them: four threads through get_report (both routes) get the single-threaded answers."""
    import threading
    import photohive_dsp_b200 as P
    from photohive_dsp_b200 import core
    imgs = [oracle.generate(k % 3, 600 + k, 640, 480) for k in range(4)]
    want = [P.get_report(im).to_json() for im in imgs]
    got = [[None] * 6 for _ in imgs]

    def work(i):
        for r in range(6):
            core._VIA_DOUBLES = bool(r & 1)  # racy on purpose: both routes interleave across the threads
            got[i][r] = P.get_report(imgs[i]).to_json()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    core._VIA_DOUBLES = False
    for i in range(4):
        assert all(g == want[i] for g in got[i])


@pytest.mark.parametrize("W,H", [(1920, 1080), (1280, 720), (2048, 1536), (640, 480)])
def test_fused_front_end_and_row_fft_launch_gives_identical_records(ctx, oracle, W, H, monkeypatch):
    """PHD_FUSED=1: front end + row FFT as one launch of role-switching persistent CTAs (fft.cu: k_front_rows).  Same
    bytes as the two separate launches, and the call reports that it did run fused."""
    torch = pytest.importorskip("torch")
    n = 24
    base = np.stack([oracle.generate(k % 3, 7000 + k, W, H) for k in range(6)])
    imgs = torch.from_numpy(np.concatenate([base] * (n // 6))).cuda()
    p = make_params()
    lay = flat_layout(p, 0)
    out = []
    for fused in ("0", "1", "1"):
        monkeypatch.setenv("PHD_FUSED", fused)
        rec = np.empty((n, lay.record_bytes), np.uint8)
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.ctypes.data)
        assert ctx.last_fused() == (fused == "1")
        out.append(rec)
    assert (out[0] == out[1]).all() and (out[1] == out[2]).all()


def test_pageable_and_pinned_host_batches_give_identical_records(ctx, oracle):
    """Pageable host input goes through the threaded pinned-slice uploader, pinned input through cudaMemcpy2DAsync."""
    torch = pytest.importorskip("torch")
    imgs = np.stack([oracle.generate(k % 3, 900 + k, 800, 600) for k in range(40)])  # > one sub-batch of 32
    p = make_params()
    lay = flat_layout(p, 0)
    pinned = torch.from_numpy(imgs).pin_memory()
    a = np.empty((40, lay.record_bytes), np.uint8)
    b = np.empty((40, lay.record_bytes), np.uint8)
    ctx.get_reports_raw(imgs.ctypes.data, 40, 800, 600, 800 * 600 * 3, p, a.ctypes.data)
    ctx.get_reports_raw(pinned.data_ptr(), 40, 800, 600, 800 * 600 * 3, p, b.ctypes.data)
    assert (a == b).all()


def test_batch_json_export_equals_single_report_json(ctx, oracle):
    """BatchReports.to_json / to_dicts carry the reference's to_json() schema (core.py:388-436) for a whole batch."""
    import json
    import photohive_dsp_b200 as P
    imgs = np.stack([oracle.generate(k % 3, 300 + k, 640, 480) for k in range(3)])
    boxes = [dict(top=0, bottom=240, left=0, right=320), dict(top=100, bottom=480, left=300, right=640)]
    barr = np.array([[[b["top"], b["bottom"], b["left"], b["right"]] for b in boxes]] * 3, np.int32)
    batch = ctx.get_reports(imgs, boxes=barr)
    dicts = batch.to_dicts(480, 640)
    for i in range(3):
        single = json.loads(P.get_report(imgs[i], salient_characters=P.set_bounding_boxes(boxes)).to_json())
        assert list(single.keys()) == list(dicts[i].keys())
        assert single == dicts[i] == json.loads(batch.to_json(i, 480, 640))


def _drop_in_on_planes(planes, W, H, params, boxes):
    from oracle import binding
    from photohive_dsp_b200 import lib as L
    lib = binding.bind_entry_points(C.CDLL(L.lib_path))
    flat = tuple(np.ascontiguousarray(p, np.float64).ravel() for p in planes)
    rp = binding.call_entry_point(lib, flat, W, H, params, boxes)
    assert rp, "get_full_report_data returned NULL"
    got = binding.unpack_full_report(rp)
    lib.free_full_report(C.byref(rp))
    return got


# palette hue of the general-input route: FP64 sums, no 2^-20 quantisation -- but the blur bins still come from the FP32 FFT
@pytest.mark.parametrize("name", ["f64_test_rgb", "f64_random", "f64_random_ties", "f64_k65535"])
def test_general_double_planes_match_reference_golden(oracle, golden, name):
    """images whose values are not k/255 -- accepted by the reference (src/image_processing.c:372-417),
    refused by round 1 -- go through the FP64 two-pass route (f64path.cu) and match the unmodified reference's outputs."""
    m = golden.meta[name]
    got = _drop_in_on_planes(golden.planes(name), m["W"], m["H"], omake(**m["params"]), m["boxes"])
    assert_report_close(got, golden_report(golden, name), name)


@pytest.mark.parametrize("W,H,kw,nbox", [
    (1920, 1080, {}, 2),
    (800, 600, dict(linked_list_size=16, coverage_thresh=0.5), 0),            # many tie groups with dropped pixels
    (752, 502, dict(downsample_rate=3, h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99), 1),
])
def test_general_double_planes_match_oracle(oracle, W, H, kw, nbox):
    rng = np.random.default_rng(W + H)
    # smooth ramps plus noise: values of every magnitude, many exactly equal channel pairs (delta == 0, max == g ...)
    y, x = np.mgrid[0:H, 0:W]
    planes = [np.clip((x / W) * 0.7 + rng.random((H, W)) * 0.3, 0, 1), np.clip((y / H) * 0.9 + rng.random((H, W)) * 0.1, 0, 1),
              np.clip(rng.random((H, W)) ** 2, 0, 1)]
    planes[1][: H // 8] = planes[0][: H // 8]      # r == g rows
    planes[2][H // 8: H // 4] = 0.0                # a zero channel: delta == max
    planes[0][-H // 8:] = planes[1][-H // 8:] = planes[2][-H // 8:] = 1.0   # max == 1 -> the 0.999999 clamps
    boxes = [dict(top=H * i // 8, bottom=H * i // 8 + H // 4, left=W * i // 8, right=W * i // 8 + W // 4) for i in range(nbox)] or None
    want = oracle.report(None, omake(**kw), boxes=boxes, nthreads=8, planes=tuple(planes))
    got = _drop_in_on_planes(planes, W, H, omake(**kw), boxes)
    assert_report_close(got, want, f"doubles {W}x{H} {kw}")
    assert len(got.palette_pct) == len(want.palette_pct)


@pytest.mark.parametrize("W,H", [(1920, 1080), (3840, 2160), (6000, 4000), (1280, 720), (2560, 1440), (1024, 768),
                                 (2048, 1536), (800, 600), (640, 480), (1008, 572), (752, 502),
                                 (7680, 4320), (12800, 2560), (2600, 11000),   # 8K; the longest row / column served
                                 # camera sizes in both orientations (compile-time plans with radices 17..21)
                                 (4032, 3024), (3024, 4032), (5472, 3648), (3648, 5472), (4000, 3000), (3000, 4000),
                                 (3264, 2448), (2448, 3264), (4608, 3456), (3456, 4608), (1600, 1200), (1200, 1600),
                                 (1280, 960), (960, 1280), (2160, 3840), (1080, 1920), (4000, 6000), (480, 600), (600, 480),
                                 (7680, 4320), (4320, 7680), (5120, 2880), (3200, 1800), (1800, 3200), (1600, 900),
                                 (1680, 1050), (1050, 1680), (2048, 1152), (2400, 1600), (2880, 1800),   # second block of plans
                                 (4096, 2160), (5184, 3456), (3000, 2000), (2704, 1520), (1520, 2704), (2592, 1944), (1944, 2592),
                                 (2011, 1511), (1511, 2011), (4030, 3020), (1031, 523),   # prime sides: Bluestein
                                 # sides beyond the shared-memory kernels: four-step transforms through HBM (13000 =
                                 # 104 x 125), and Bluestein around them for a prime side (13003)
                                 (13000, 2600), (2600, 13000), (13003, 2602), (2602, 13003)])
def test_power_spectrum_against_float64_fft(ctx, oracle, W, H, record_property):
    """Every compile-time FFT plan (and the runtime-radix / Bluestein shapes) against numpy's float64 rfft2 of the same
    exact gray numerators, element by element (bounds and what is reported: tests/parity.py, "FFT magnitudes")."""
    img = oracle.generate(0, 31 + W, W, H)
    pw = ctx.debug_power_spectrum(img).astype(np.float64)
    i64 = img.astype(np.int64)
    gnum = (299 * i64[:, :, 0] + 587 * i64[:, :, 1] + 114 * i64[:, :, 2]) - 127500
    ref = np.abs(np.fft.rfft2(gnum.astype(np.float64) / 255000.0)) ** 2
    assert pw.shape == ref.shape
    rep = fft_error_report(pw, ref)
    record_property("fft_error", rep)
    print(f"FFT {W}x{H}: |X| rel err median {rep['median']:.2e} p99 {rep['p99']:.2e} max {rep['max']:.2e} "
          f"(max over the {rep['n_significant']} coefficients >= 1e-3 of the mean power: {rep['max_significant']:.2e}); "
          f"p<1 threshold flips: {rep['threshold_flips']} of {rep['n']}")
    err = np.abs(pw - ref) / (ref + ref.mean())
    assert err.max() < 1e-4, (W, H, float(err.max()), np.unravel_index(err.argmax(), err.shape))
    # power relative error = 2 x magnitude relative error (prime sides go through Bluestein: 2.1e-5 on |X| at 2011x1511,
    # where the O(p^2) pass of round 1 had 6.5e-5)
    assert 2 * rep["max_significant"] < 1e-4, rep
    assert rep["median"] < 2e-6 and rep["p99"] < 1e-4, rep


# ---- full-size, size-independent properties --------------------------------------------------------
@pytest.mark.parametrize("W,H", [(3840, 2160), (6000, 4000), (7680, 4320)])
def test_full_size_properties(ctx, oracle, W, H):
    """BASELINE configs 2 and 4 (4K, 24 MP): invariants that need no CPU run of the whole pipeline."""
    img = oracle.generate(1, 2024, W, H)
    P = W * H
    p = make_params()
    b = ctx.get_reports(img[None], params=p)
    n = int(b.palette_n[0])
    # palette: every pixel is counted once or dropped by the tie path; parents are distinct groups
    assert abs(b.palette_pct[0, :n].sum() - (1 - int(b.dropped_pixels[0]) / P)) < 1e-12
    assert len(set(b.parent_ids[0, :n].tolist())) == n
    counts = ctx.debug_group_counts(img, p)
    assert counts.sum() == P
    # exact channel statistics from integer sums
    k = img.reshape(-1, 3).astype(np.int64)
    s1, s2 = k.sum(0), (k * k).sum(0)   # exact integers; numpy's float mean over 8-24 M values drifts by 1e-11
    assert np.all(rel_err(b.rgb_stats[0, :3], s1 / 255.0 / P) < 1e-13)
    var = (s2 * P - s1 * s1).astype(np.float64) / (float(P) * P * 65025.0)
    assert np.all(rel_err(b.rgb_stats[0, 3:], np.sqrt(var)) < 1e-12)
    # Parseval on the hand-written 2-D FFT: sum over the full spectrum of |X|^2 == W*H * sum x^2
    pw = ctx.debug_power_spectrum(img).astype(np.float64)
    i64 = img.astype(np.int64)
    gnum = (299 * i64[:, :, 0] + 587 * i64[:, :, 1] + 114 * i64[:, :, 2]) - 127500
    energy = float(np.sum(gnum.astype(np.float64) ** 2)) / 255000.0 ** 2 * P
    wts = np.full(W // 2 + 1, 2.0)
    wts[0] = 1.0
    if W % 2 == 0:
        wts[-1] = 1.0
    assert abs(np.sum(pw * wts[None, :]) - energy) / energy < 1e-5
    # DC bin is the plain sum
    assert abs(np.sqrt(pw[0, 0]) - abs(gnum.sum()) / 255000.0) / (abs(gnum.sum()) / 255000.0) < 1e-5
    # bins: empty-bin pattern comes from the (exact) bin map; values are normalised to [0, 1]
    _, cnt = ctx.debug_bin_map(W, H)
    assert np.array_equal(b.blur_bins[0] == 0, cnt == 0) or np.all(b.blur_bins[0][cnt == 0] == 0)
    assert b.blur_bins[0].max() <= 1.0 + 1e-9 and b.blur_bins[0].min() >= 0
