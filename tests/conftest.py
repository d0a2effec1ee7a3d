import json
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement (TEST infrastructure).  Built on demand with gcc."""
    from oracle import binding
    binding.build(ref=False)
    return binding.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled at -O0 (only where /root/reference or a prebuilt _ref exists)."""
    from oracle import binding
    if not binding.Reference.available(0):
        if not os.path.isdir("/root/reference/src"):
            pytest.skip("compiled reference not available here")
        binding.build(ref=True)
    return binding.Reference(0)


class Golden:
    def __init__(self):
        z = np.load(GOLDEN)
        self.z = z
        self.meta = json.loads(bytes(z["meta"]).decode())

    def names(self, max_pixels=None):
        return [n for n, m in self.meta.items() if "planes" not in m and (max_pixels is None or m["W"] * m["H"] <= max_pixels)]

    def plane_names(self):
        return [n for n, m in self.meta.items() if "planes" in m]

    def image(self, oracle, name):
        m = self.meta[name]
        img = oracle.generate(m["kind"], m["seed"], m["W"], m["H"])
        assert zlib.crc32(img.tobytes()) == m["crc32"], "synthetic generator no longer reproduces the fixture input"
        return img

    def planes(self, name):
        """Three float64 planes of a general-input (not k/255) fixture, regenerated and CRC-guarded like the images."""
        from parity import planes_for
        m = self.meta[name]
        planes = planes_for(m["planes"], m["seed"], m["W"], m["H"])
        assert zlib.crc32(np.stack(planes).tobytes()) == m["crc32"], "plane generator no longer reproduces the fixture input"
        return planes

    def field(self, name, key):
        k = f"{name}/{key}"
        return self.z[k] if k in self.z.files else None


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI (fails loudly when the library or the device is missing)."""
    from photohive_dsp_b200.batch import Context
    c = Context(0)
    yield c
    c.close()
