"""Host-side export helpers (no GPU): the vectorised palette colour conversion equals the scalar one the reference's
Report uses (utils.py:8-28), including sector edges."""
import numpy as np

from photohive_dsp_b200.batch import BatchReports
from photohive_dsp_b200.utils import hsv_to_rgb


def test_palette_rgb_matches_scalar_hsv_to_rgb():
    rng = np.random.default_rng(0)
    hsv = np.stack([rng.uniform(0, 360, (3, 200)), rng.uniform(0, 1, (3, 200)), rng.uniform(0, 1, (3, 200))], -1)
    edges = [(0, 0, 0), (60, .5, .5), (120, 1, 1), (180, .25, .75), (240, .9, .1), (300, .3, .9), (359.999999, 1, 1),
             (59.99999999, .7, .7), (0, 0, 0.999999), (180, 0.999999, 0.999999)]
    hsv[0, :len(edges)] = edges
    b = BatchReports.__new__(BatchReports)
    b.palette_hsv = hsv
    want = np.array([[hsv_to_rgb(*hsv[i, j]) for j in range(hsv.shape[1])] for i in range(hsv.shape[0])])
    assert np.array_equal(b.palette_rgb(), want)
