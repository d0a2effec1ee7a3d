"""Device-resident throughput of the other BASELINE.json configs (parity-test cases, not bench lines):
config 2 (4K + 4 boxes), config 4 (24 MP), config 5 (fine palette on 1080p).  Prints one line each."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from tools.synth import Generator  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)


def run(name, W, H, n, params, boxes=None, reps=3):
    gen = Generator(W, H, dev)
    imgs = gen.batch(n)
    nb = 0 if boxes is None else boxes.shape[1]
    lay = flat_layout(params, nb)
    rec = torch.empty((n, lay.record_bytes), dtype=torch.uint8, device=dev)
    barr = None if boxes is None else np.ascontiguousarray(boxes, np.int32)
    bkw = {} if barr is None else dict(boxes_ptr=barr.ctypes.data, max_boxes=nb)
    for _ in range(2):
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, params, rec.data_ptr(), **bkw)
    tot = 0.0
    stages = {}
    for _ in range(reps):
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, params, rec.data_ptr(), **bkw)
        ms, _ = ctx.last_timing()
        tot += ms["total"]
        for k, v in ms.items():
            stages[k] = stages.get(k, 0.0) + v / reps
    print(f"{name}: {n * reps / (tot / 1000.0):.0f} images/s  ({tot / reps / n * 1000:.1f} us/image)  " +
          "  ".join(f"{k}={v / n * 1000:.2f}" for k, v in stages.items() if k != "total"), flush=True)
    del imgs, rec
    torch.cuda.empty_cache()


p = make_params()
W, H = 3840, 2160
boxes = np.array([[[H * i // 8, H * i // 8 + H // 4, W * i // 8, W * i // 8 + W // 4] for i in range(4)]] * 128, np.int32)
run("config2 4K + 4 boxes x128", W, H, 128, p, boxes)
run("config4 24MP x64", 6000, 4000, 64, p)
if "--only-big" in sys.argv:
    sys.exit(0)
run("config5 fine palette 1080p x256", 1920, 1080, 256, make_params(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99))
run("config3 1080p x512", 1920, 1080, 512, p)
if "--generic" in sys.argv:  # other shapes: camera sizes with compile-time plans, and the runtime-radix kernels
    run("12 MP phone 4032x3024 x32", 4032, 3024, 32, p)
    run("20 MP 5472x3648 x16", 5472, 3648, 16, p)
    run("720p 1280x720 x256", 1280, 720, 256, p)
    run("portrait 1080x1920 x256 (rows: runtime-radix kernel, width not a multiple of 16)", 1080, 1920, 256, p)
    run("portrait 3024x4032 x32", 3024, 4032, 32, p)
    run("8K 7680x4320 x16 (runtime-radix FFT kernels)", 7680, 4320, 16, p)
    run("4030x3020 x8 (runtime-radix FFT kernels, 2*5*13*31 x 4*5*151)", 4030, 3020, 8, p)
