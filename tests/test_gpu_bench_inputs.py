"""The bench's inputs are what they claim to be: tools/synth.py (images generated directly in HBM) reproduces
oracle.generate() byte for byte, and record 0 of bench.py's workload is the committed golden case g0_1080p
(tests/golden/reference_golden.npz, written by the UNMODIFIED reference).  Run with -m gpu."""
import numpy as np
import pytest

from parity import assert_report_close, golden_report, report_from_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(1920, 1080), (640, 480)])
def test_device_generator_equals_oracle_generator(oracle, W, H):
    import torch
    from tools.synth import Generator
    gen = Generator(W, H, torch.device("cuda", 0))
    for kind, seed in ((0, 12345), (1, 12346), (0, 99), (1, 7)):
        got = gen.image(kind, seed).cpu().numpy()
        want = oracle.generate(kind, seed, W, H)
        assert got.shape == want.shape and got.dtype == want.dtype
        assert np.array_equal(got, want), f"G{kind} seed {seed}: {np.count_nonzero(got != want)} bytes differ"
    batch = gen.batch(4, 12345).cpu().numpy()
    for i in range(4):
        assert np.array_equal(batch[i], oracle.generate(i % 2, 12345 + i, W, H))


def test_bench_record_zero_is_the_golden_case(ctx, oracle, golden):
    """bench.py's image 0 (G0, seed 12345, 1920x1080, defaults) is golden `g0_1080p`, image 1 shares the generator and
    seed rule of `g1_1080p` (seed 12346 there is not a fixture, so it is checked against the oracle)."""
    import torch
    from tools.synth import Generator
    import bench
    assert (bench.W, bench.H, bench.FIRST_SEED) == (1920, 1080, 12345)
    m = golden.meta["g0_1080p"]
    assert (m["kind"], m["seed"], m["W"], m["H"], m["params"], m["boxes"]) == (0, 12345, 1920, 1080, {}, None)
    images = Generator(1920, 1080, torch.device("cuda", 0)).batch(2, bench.FIRST_SEED)
    b = ctx.get_reports(images)
    assert_report_close(report_from_batch(b, 0), golden_report(golden, "g0_1080p"), "bench image 0 vs golden g0_1080p")
    from oracle.binding import make_params as omake
    want = oracle.report(oracle.generate(1, 12346, 1920, 1080), omake(), nthreads=8)
    assert_report_close(report_from_batch(b, 1), want, "bench image 1 vs oracle")
    bench.check_record_zero(b.raw[0], b.layout)  # the same check bench.py runs outside its timed region
