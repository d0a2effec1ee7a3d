"""End-to-end throughput of the Python batch call on a PAGEABLE numpy batch (what `get_reports(np_array)` users see)
next to the same call on pinned memory: python tools/bench_pageable.py [n_images]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from tools.synth import Generator  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, H = 1920, 1080
dev = torch.device("cuda", 0)
ctx = Context(0)
params = make_params()
lay = flat_layout(params, 0)
imgs = Generator(W, H, dev).batch(n).cpu()
pageable = imgs.numpy().copy()
pinned = imgs.pin_memory()
rec = np.empty((n, lay.record_bytes), np.uint8)
rec2 = np.empty((n, lay.record_bytes), np.uint8)
for name, ptr in (("pageable numpy", pageable.ctypes.data), ("pinned", pinned.data_ptr())):
    out = rec if name.startswith("pageable") else rec2
    for _ in range(2):
        ctx.get_reports_raw(ptr, n, W, H, W * H * 3, params, out.ctypes.data)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ctx.get_reports_raw(ptr, n, W, H, W * H * 3, params, out.ctypes.data)
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {n / dt:.0f} images/s ({n * W * H * 3 / dt / 1e9:.1f} GB/s of input)", flush=True)
print("records identical:", bool((rec == rec2).all()))
