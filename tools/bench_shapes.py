"""Device-resident throughput of arbitrary shapes: python tools/bench_shapes.py WxH[xN] ...  (N images, default 16)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from tools.synth import Generator  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)
p = make_params()
for spec in sys.argv[1:]:
    parts = [int(x) for x in spec.split("x")]
    W, H, n = parts[0], parts[1], (parts[2] if len(parts) > 2 else 16)
    imgs = Generator(W, H, dev).batch(n)
    lay = flat_layout(p, 0)
    rec = torch.empty((n, lay.record_bytes), dtype=torch.uint8, device=dev)
    for _ in range(2):
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.data_ptr())
    tot, stages, reps = 0.0, {}, 3
    for _ in range(reps):
        ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.data_ptr())
        ms, _ = ctx.last_timing()
        tot += ms["total"]
        for k, v in ms.items():
            stages[k] = stages.get(k, 0.0) + v / reps
    print(f"{W}x{H} x{n}: {n * reps / (tot / 1000.0):.0f} images/s  ({tot / reps / n * 1000:.1f} us/image)  " +
          "  ".join(f"{k}={v / n * 1000:.1f}" for k, v in stages.items() if k != "total"), flush=True)
    del imgs, rec
    torch.cuda.empty_cache()
