"""Tiny workload for compute-sanitizer (racecheck / memcheck) where the tool is available (it is closed on the pool this
round ran on; tools/check_determinism.py is the stand-in): a few images through every kernel family.
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py
Covers the compile-time FFT plans (640x480 by default, W H as arguments), boxes (sharpness), the tie path and, with
`generic`, a shape that takes the runtime-radix FFT kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
import torch  # noqa: E402

from tools.synth import Generator  # noqa: E402

args = [a for a in sys.argv[1:] if a != "generic"]
W, H = (int(args[0]), int(args[1])) if len(args) >= 2 else (640, 480)
if "generic" in sys.argv:
    W, H = 646, 456
n = 3
ctx = Context(0)
p = make_params()
imgs = Generator(W, H, torch.device('cuda', 0)).batch(n)
boxes = np.array([[[0, H // 2, 0, W // 2], [H // 4, H, W // 3, W]]] * n, np.int32)
b = ctx.get_reports(imgs, boxes=boxes, params=p)
print("ok", W, H, b.palette_n.tolist())
