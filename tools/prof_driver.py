"""Small fixed workload for ncu: N device-resident images (G0/G1; 1080p unless W H are given), default parameters, two calls.
  python tools/prof_driver.py [N=32] [W H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from photohive_dsp_b200.batch import Context, flat_layout, make_params  # noqa: E402
from tools.synth import Generator  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
dev = torch.device("cuda", 0)
ctx = Context(0)
p = make_params()
lay = flat_layout(p, 0)
imgs = Generator(W, H, dev).batch(n)
rec = torch.empty((n, lay.record_bytes), dtype=torch.uint8, device=dev)
for _ in range(2):
    ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.data_ptr())
torch.cuda.synchronize()
print(ctx.last_timing())
