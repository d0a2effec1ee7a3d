"""Re-runs the device pipeline on the same resident batch and counts records whose bytes change from run to run (there
must be none: every accumulator is an integer and every shared-memory hand-over is fenced).  A race in the row FFT's
thread-group barriers was found with this (sequences of two groups shared a buffer region): python tools/check_determinism.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from photohive_dsp_b200.batch import Context, flat_layout, make_params
from tools.synth import Generator
dev = torch.device("cuda", 0)
ctx = Context(0)
p = make_params()
lay = flat_layout(p, 0)
def run(W, H, n, reps=6, p=p, boxes=False):
    imgs_d = Generator(W, H, dev).batch(n)
    nb = 2 if boxes else 0
    lay = flat_layout(p, nb)
    barr = np.ascontiguousarray([[[0, H // 2, 0, W // 2], [H // 4, H, W // 3, W]]] * n, np.int32)
    bkw = dict(boxes_ptr=barr.ctypes.data, max_boxes=nb) if boxes else {}
    outs = []
    for r in range(reps):
        c = np.empty((n, lay.record_bytes), np.uint8)
        ctx.get_reports_raw(imgs_d.data_ptr(), n, W, H, W*H*3, p, c.ctypes.data, **bkw)
        outs.append(c)
    bad = 0
    first_off = None
    for c in outs[1:]:
        d = (c != outs[0])
        bad += int(d.any(axis=1).sum())
        if d.any() and first_off is None:
            i = int(np.nonzero(d.any(axis=1))[0][0]); offs = np.nonzero(d[i])[0]
            first_off = (i, int(offs[0]), int(offs[-1]), len(offs))
    print(W, H, n, "records differing from run 0 over", reps - 1, "reruns:", bad, first_off, flush=True)
for W, H, n in ((800, 600, 40), (800, 600, 400), (1920, 1080, 64), (1280, 720, 64), (640, 480, 64), (1024, 768, 64), (752, 502, 32)):
    run(W, H, n)
run(3840, 2160, 16, reps=4, boxes=True)
run(6000, 4000, 6, reps=4)
run(4032, 3024, 4, reps=4)
run(2560, 1440, 16, reps=4)
run(2048, 1536, 16, reps=4)
run(1920, 1080, 32, reps=4, p=make_params(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99))
run(1920, 1080, 32, reps=4, p=make_params(downsample_rate=2), boxes=True)
for W, H, n in ((4032, 3024, 8), (3024, 4032, 8), (5472, 3648, 6), (3648, 5472, 6), (3264, 2448, 8), (4608, 3456, 6),
                (1600, 1200, 16), (960, 1280, 16), (2160, 3840, 8), (4000, 6000, 4)):
    run(W, H, n, reps=4)
