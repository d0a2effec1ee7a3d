"""Developer smoke check on a GPU box: prints GPU-vs-oracle differences field by field.
Usage: python tools/quick_check.py [WxH ...]   (the oracle is TEST infrastructure; this script is not product)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.binding import Oracle, make_params as omake  # noqa: E402
from photohive_dsp_b200.batch import Context, make_params  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))) if a.size else 0.0


def compare(tag, g, i, o):
    n = int(g.palette_n[i])
    print(f"[{tag}] N gpu={n} oracle={len(o.palette_pct)} ties={int(g.tie_groups[i])}/{o.extra['tie_groups']} "
          f"dropped={int(g.dropped_pixels[i])}/{o.extra['dropped_pixels']}")
    print("   rgb_stats rel", rel(g.rgb_stats[i], o.rgb_stats), " sat rel", rel(g.average_saturation[i], o.average_saturation))
    if n == len(o.palette_pct):
        print("   parents equal", bool(np.array_equal(g.parent_ids[i, :n], o.extra["parent_ids"])),
              " pct maxabs", float(np.max(np.abs(g.palette_pct[i, :n] - o.palette_pct))),
              " hsv maxabs", np.max(np.abs(g.palette_hsv[i, :n] - o.palette_hsv), axis=0))
    print("   bins rel", rel(g.blur_bins[i], o.blur_bins), " maxabs", float(np.max(np.abs(g.blur_bins[i] - o.blur_bins))),
          " abs/rbs", int(g.angle_bin_size[i]), o.angle_bin_size, int(g.radius_bin_size[i]), o.radius_bin_size)
    print("   vec", g.blur_vec_angle[i].tolist(), o.blur_vec_angle.tolist(), g.blur_vec_mag[i].tolist(), o.blur_vec_mag.tolist())
    if o.sharpness is not None:
        print("   sharp", g.sharpness[i], o.sharpness, "rel", rel(g.sharpness[i], o.sharpness))


def main():
    sizes = [(1920, 1080)]
    if len(sys.argv) > 1:
        sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    orc = Oracle()
    ctx = Context(0)
    p, po = make_params(), omake()
    t = time.time()
    sw = ctx.debug_group_sweep(p)
    print("group sweep gpu %.2fs" % (time.time() - t))
    t = time.time()
    swo = orc.group_sweep(po)
    print("group sweep oracle %.2fs; mismatches: %d" % (time.time() - t, int(np.count_nonzero(sw != swo))))
    for (W, H) in sizes:
        m, c = ctx.debug_bin_map(W, H)
        mo, co = orc.bin_map(W, H)
        print(f"bin map {W}x{H}: id mismatches {int(np.count_nonzero(m != mo))}, count mismatches {int(np.count_nonzero(c != co))}")
        imgs = np.stack([orc.generate(k, 12345 + k, W, H) for k in (1, 0, 2)])
        boxes = np.array([[[H * i // 8, H * i // 8 + H // 4, W * i // 8, W * i // 8 + W // 4] for i in range(4)]] * 3, np.int32)
        t = time.time()
        g = ctx.get_reports(imgs, boxes=boxes, params=p)
        print(f"gpu batch of 3 {W}x{H}: %.3fs" % (time.time() - t), ctx.last_timing())
        t = time.time()
        g = ctx.get_reports(imgs, boxes=boxes, params=p)
        print("second call: %.3fs" % (time.time() - t), ctx.last_timing())
        for i, k in enumerate((1, 0, 2)):
            bl = [dict(top=int(b[0]), bottom=int(b[1]), left=int(b[2]), right=int(b[3])) for b in boxes[i]]
            o = orc.report(imgs[i], po, boxes=bl, nthreads=8, want_intermediates=(i == 0))
            compare(f"G{k} {W}x{H}", g, i, o)
            if i == 0:
                pw = ctx.debug_power_spectrum(imgs[i]).astype(np.float64)
                # oracle power is of gray-avg; ours of gray-0.5: identical except [0,0]
                po_ = o.extra["power"].reshape(H, W // 2 + 1).copy()
                pw[0, 0] = po_[0, 0]
                rms = np.sqrt(np.mean(po_))
                print("   power: max |d|/rms(power) ", float(np.max(np.abs(pw - po_)) / np.mean(po_)),
                      " median rel", float(np.median(np.abs(pw - po_) / np.maximum(po_, 1e-30))), "rms", rms)
        # fine palette + downsample
        for kw in (dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99), dict(downsample_rate=5)):
            g2 = ctx.get_reports(imgs[:2], params=make_params(**kw))
            for i in range(2):
                o = orc.report(imgs[i], omake(**kw), nthreads=8)
                compare(f"{kw} img{i}", g2, i, o)


if __name__ == "__main__":
    main()
