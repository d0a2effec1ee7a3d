import os, sys
sys.path.insert(0, os.getcwd())
import torch
from photohive_dsp_b200.batch import Context, flat_layout, make_params
from tools.synth import Generator
n = 256
W, H = 1920, 1080
dev = torch.device("cuda", 0)
ctx = Context(0)
p = make_params(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99)
lay = flat_layout(p, 0)
imgs = Generator(W, H, dev).batch(n)
rec = torch.empty((n, lay.record_bytes), dtype=torch.uint8, device=dev)
for _ in range(2):
    ctx.get_reports_raw(imgs.data_ptr(), n, W, H, W * H * 3, p, rec.data_ptr())
torch.cuda.synchronize()
print(ctx.last_timing())
