"""Per-stage device time on batches of ONE generator kind (G0 noise / G1 gradient+noise) and on a smooth image."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from tools.synth import Generator  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)
W, H, n = 1920, 1080, 256
p = make_params()
lay = flat_layout(p, 0)
gen = Generator(W, H, dev)
rec = torch.empty((n, lay.record_bytes), dtype=torch.uint8, device=dev)
for name in ("G0", "G1", "smooth"):
    imgs = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(n):
        if name == "G0":
            gen.image(0, 100 + i, imgs[i])
        elif name == "G1":
            gen.image(1, 100 + i, imgs[i])
        else:
            imgs[i].copy_((gen.base + (i & 15)).to(torch.uint8) & 255)
    for _ in range(2):
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.data_ptr())
    ms, _ = ctx.last_timing()
    print(name, "  ".join(f"{k}={v / n * 1000:.2f}" for k, v in ms.items()), flush=True)
    del imgs
