"""Copy-only probe of the host -> device path: what the PLATFORM gives N GPUs copying from pinned host memory at the
same time, with no kernel of ours involved.  It answers VERDICT r1's question whether the end-to-end scaling beyond two
GPUs (SCALE_r01: 1.00 / 1.00 / 0.53 / 0.43) is the host or the library's staging scheme.

    python tools/h2d_probe.py [--gpus 1,2,4,8] [--mb 1024] [--reps 8]      (one process, one host thread per GPU)

Per GPU count it reports the aggregate and per-GPU GB/s of (a) plain 1-D cudaMemcpyAsync of one large pinned buffer
and (b) the library's own pattern -- cudaMemcpy2DAsync of 32 images of 1920x1080x3 bytes per call
(pipeline.cu: stage_copy) -- plus the images/s those rates allow at 6,220,800 bytes per image.  torch is used for the
pinned / device allocations and streams only."""
import argparse
import ctypes as C
import json
import threading
import time

import torch

IMG = 1920 * 1080 * 3


def cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    raise RuntimeError("libcudart not found")


def run(devs, mb, reps, rt):
    n = len(devs)
    nbytes = mb << 20
    nimg = nbytes // IMG
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in devs]
    devb = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]
    for hb in host:
        hb.random_(0, 255)
    out = {}
    for mode in ("memcpy1d", "memcpy2d_32_images"):
        start = threading.Barrier(n + 1)
        done = threading.Barrier(n + 1)
        per = [0.0] * n

        def worker(i):
            d = devs[i]
            torch.cuda.set_device(d)
            st = streams[i]
            sp = C.c_void_p(st.cuda_stream)
            for _ in range(2):  # warm-up
                devb[i].copy_(host[i], non_blocking=True)
            torch.cuda.synchronize(d)
            start.wait()
            t0 = time.perf_counter()
            with torch.cuda.stream(st):
                for _ in range(reps):
                    if mode == "memcpy1d":
                        devb[i].copy_(host[i], non_blocking=True)
                    else:
                        for first in range(0, nimg - 31, 32):
                            rc = rt.cudaMemcpy2DAsync(C.c_void_p(devb[i].data_ptr() + first * IMG), C.c_size_t(IMG),
                                                      C.c_void_p(host[i].data_ptr() + first * IMG), C.c_size_t(IMG),
                                                      C.c_size_t(IMG), C.c_size_t(32), C.c_int(1), sp)
                            assert rc == 0, rc
            st.synchronize()
            per[i] = time.perf_counter() - t0
            done.wait()

        ts = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
        for t in ts:
            t.start()
        start.wait()
        t0 = time.perf_counter()
        done.wait()
        wall = time.perf_counter() - t0
        for t in ts:
            t.join()
        moved = (nbytes if mode == "memcpy1d" else (nimg // 32) * 32 * IMG) * reps
        out[mode] = {"aggregate_GBps": n * moved / wall / 1e9, "per_gpu_GBps": [moved / p / 1e9 for p in per],
                     "images_per_s_at_6.2MB": n * moved / wall / IMG}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    rt = cudart()
    have = torch.cuda.device_count()
    res = {"device": torch.cuda.get_device_name(0), "visible_gpus": have, "pinned_buffer_MiB": a.mb, "reps": a.reps, "runs": {}}
    for g in [int(x) for x in a.gpus.split(",")]:
        if g > have:
            continue
        res["runs"][str(g)] = run(list(range(g)), a.mb, a.reps, rt)
        r = res["runs"][str(g)]
        print(f"{g} GPU(s): 1-D {r['memcpy1d']['aggregate_GBps']:.1f} GB/s aggregate "
              f"({min(r['memcpy1d']['per_gpu_GBps']):.1f}..{max(r['memcpy1d']['per_gpu_GBps']):.1f} per GPU), "
              f"library pattern {r['memcpy2d_32_images']['aggregate_GBps']:.1f} GB/s = "
              f"{r['memcpy2d_32_images']['images_per_s_at_6.2MB']:.0f} images/s", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
