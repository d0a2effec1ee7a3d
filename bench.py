#!/usr/bin/env python
"""bench.py -- images/s of the full get_report() pipeline on synthetic RGB batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {2,3,4,5}]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the hot path over one batch of synthetic images per GPU.  The default workload is BASELINE
config 3 (1920x1080, image i = G0/G1 alternating with seed 12345+i, defaults h18 s2 v3); --config 2 / 4 / 5 select the
other BASELINE.json configurations (4K with four salient boxes, 24 MP, fine palette at 1080p) with the same JSON
schema.  Images are independent, so N GPUs run N shards with no data-path collective (weak scaling: the per-GPU batch
is fixed); rank 0 prints ONE JSON line.  `value` is measured with the inputs resident in HBM; `e2e` goes through the same
C-ABI call with pinned HOST buffers, H2D and D2H inside the timed region, and under torchrun every rank's device writes
its records into ITS SLICE OF ONE SHARED PINNED HOST ARRAY (the host-side gather of SURVEY.md section 8e: rank 0 holds
every record when the closing barrier of the timed region returns).  Launched WITHOUT torchrun and --gpus N > 1 the
same is done by one process through phd_get_reports_u8_multi (one host thread per GPU).  `cpu_baseline` /
`--impl reference` time the reference's own CPU code (oracle/_ref, built from the unmodified sources) on this box's
host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080   # the default workload (config 3); other configs carry their own shape
FIRST_SEED = 12345
METRIC = "images/sec full get_report (1080p RGB)"
FINE = dict(h_partitions=36, s_partitions=4, v_partitions=6, coverage_thresh=0.99)


def boxes_for(w, h, n=4):
    """The four salient boxes of config 2 (W/4 x H/4 each, i.e. 960x540 at 4K) as
    (top, bottom, left, right) -- the same formula the golden fixtures use (tests/golden/make_golden.py)."""
    return [(h * i // 8, h * i // 8 + h // 4, w * i // 8, w * i // 8 + w // 4) for i in range(n)]


CONFIGS = {
    3: dict(w=1920, h=1080, batch=4096, e2e=512, params={}, nbox=0, metric=METRIC,
            name="BASELINE config 3: {batch} x 1920x1080 packed 8-bit RGB per GPU per step, image i = G0/G1 alternating "
                 "with seed {seed}+i, defaults h18 s2 v3 cov .95, 72x40 blur bins, no boxes"),
    2: dict(w=3840, h=2160, batch=1024, e2e=128, params={}, nbox=4,
            metric="images/sec full get_report (3840x2160 RGB, 4 salient boxes)",
            name="BASELINE config 2: {batch} x 3840x2160 packed 8-bit RGB per GPU per step, each with 4 salient "
                 "Crop_Boundaries of 960x540 (sharpness path), image i = G0/G1 alternating with seed {seed}+i, defaults"),
    4: dict(w=6000, h=4000, batch=512, e2e=32, params={}, nbox=0,
            metric="images/sec full get_report (6000x4000 RGB, 24 MP)",
            name="BASELINE config 4: {batch} x 6000x4000 (24 MP) packed 8-bit RGB per GPU per step, blur-profile FFT "
                 "stress, image i = G0/G1 alternating with seed {seed}+i, defaults"),
    5: dict(w=1920, h=1080, batch=4096, e2e=512, params=FINE, nbox=0,
            metric="images/sec full get_report (1080p RGB, fine palette h36 s4 v6 cov .99)",
            name="BASELINE config 5: {batch} x 1920x1080 packed 8-bit RGB per GPU per step, fine palette h=36 s=4 v=6 "
                 "coverage_thresh=0.99 (histogram contention stress), image i = G0/G1 alternating with seed {seed}+i"),
}


def algo_bytes(w=W, h=H, box_area=0):
    """read RGB once + write and read the FP32 half spectrum once (+ boxes)."""
    return 3 * w * h + 2 * 8 * h * (w // 2 + 1) + 3 * box_area


def kernel_bytes(w=W, h=H, box_area=0, fused=False):
    """Algorithmic bytes per image of each stage (its share of the figure above).  With the fused
    front-end + row launch the RGB is read once for both."""
    fw = w // 2 + 1
    kb = {"frontend": 3 * w * h, "fft_rows": 3 * w * h + 8 * h * fw, "fft_cols_blur": 8 * h * fw + 2 * h * fw,
          "sharpness": 3 * box_area}
    if fused:
        kb["frontend"] = 3 * w * h + 8 * h * fw
        kb["fft_rows"] = 0
    return kb


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def check_record_zero(raw_record, lay):
    """Outside the timed region: record 0 of the default workload (G0, seed 12345, 1920x1080, default parameters) must
    equal the committed golden case `g0_1080p` -- outputs of the UNMODIFIED reference (tests/golden/make_golden.py) --
    within the tolerances of tests/parity.py.  Raises AssertionError otherwise."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Golden
    from parity import assert_report_close, golden_report, report_from_batch
    from photohive_dsp_b200.batch import view_records
    g = Golden()
    m = g.meta["g0_1080p"]
    assert (m["kind"], m["seed"], m["W"], m["H"]) == (0, FIRST_SEED, W, H) and not m["params"] and not m["boxes"]
    rec = np.ascontiguousarray(np.asarray(raw_record, np.uint8).reshape(1, -1))
    assert_report_close(report_from_batch(view_records(rec, lay), 0), golden_report(g, "g0_1080p"), "bench record 0")
    return "record 0 equals golden g0_1080p (unmodified reference) within tests/parity.py tolerances"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, opt, seeds, cfg_id = args
    cfg = CONFIGS[cfg_id]
    w, h = cfg["w"], cfg["h"]
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # the reference prints ~20 timing lines per call
    from oracle import binding
    orc = binding.Oracle()
    impl = binding.Reference(opt) if kind == "reference" else orc
    boxes = [dict(top=t, bottom=b, left=l, right=r) for t, b, l, r in boxes_for(w, h, cfg["nbox"])] or None
    t_total = 0.0
    for i, seed in seeds:
        img = orc.generate(i % 2, seed, w, h)
        planes = binding.planes_from_u8(img)
        t0 = time.perf_counter()
        if kind == "reference":
            r = impl.report(None, binding.make_params(**cfg["params"]), boxes=boxes,
                            planes=tuple(p.reshape(h, w) for p in planes))
        else:
            r = impl.report(img, binding.make_params(**cfg["params"]), boxes=boxes, nthreads=1)
        t_total += time.perf_counter() - t0
        assert r is not None
    return t_total


def cpu_arm(images_per_core: int = 2, max_procs: int = 64, cfg_id: int = 3):
    """images/s of the reference CPU path with one process per core (FFT threads = 1 each would be unfair to
    nobody: the reference asks FFTW for nproc threads per call; one process per core keeps every core busy)."""
    import multiprocessing as mp
    from oracle import binding
    binding.build(ref=True)
    kind, opt = ("reference", 2) if binding.Reference.available(2) else ("port", 0)
    cores = min(os.cpu_count() or 1, max_procs)
    os.environ["OMP_NUM_THREADS"] = "1"
    jobs = [(kind, opt, [(c * images_per_core + j, FIRST_SEED + c * images_per_core + j) for j in range(images_per_core)],
             cfg_id) for c in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        busy = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    n = cores * images_per_core
    # throughput from the per-process busy time (excludes interpreter start-up and image generation)
    value = n / max(busy) if busy else 0.0
    return dict(value=value, unit="images/s", cores=cores, kind=kind,
                sample=f"{n} images of the workload ({images_per_core} per process, one process per core, "
                       f"{'unmodified reference sources at -O2 + FFT stand-in (FFTW absent)' if kind == 'reference' else 'oracle port'}); "
                       f"wall {wall:.1f}s incl. start-up"), wall


# ------------------------------------------------------------------------------------------------
def clocks_sampler_start(path):
    try:
        return subprocess.Popen(
            ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
             "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def clocks_summary(path, gpu_index):
    sm, mx, reasons = [], 0.0, set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) != gpu_index:
                continue
            sm.append(float(f[1])); mx = max(mx, float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    top = sorted(sm)[len(sm) // 2:]  # samples under load are the upper half
    return {"sm_mhz": float(np.median(top)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def shared_pinned_array(name, nbytes, create):
    """One host array every rank of the node maps (POSIX shared memory) and registers with CUDA as pinned memory: each
    rank's device writes its records straight into its slice -- the host-side gather of SURVEY.md section 8(e)."""
    import torch
    path = os.path.join("/dev/shm", name)
    if create:
        with open(path, "wb") as f:
            f.truncate(nbytes)
    t = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
    pinned = (int(rc) == 0)
    return t, path, pinned


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json configuration (default 3)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PHD_BENCH_BATCH", 0)), help="images per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    w, h = cfg["w"], cfg["h"]
    batch = args.batch or cfg["batch"]
    eb = min(args.e2e_batch or cfg["e2e"], batch)

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    in_process = world == 1 and args.gpus > 1   # one process drives every GPU through phd_get_reports_u8_multi
    workload = cfg["name"].format(batch=batch, seed=FIRST_SEED)

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(args.steps, 1)
        vals = []
        base = None
        for _ in range(max(args.warmup, 0) and 1):  # one warm-up pass is enough to page the libraries in
            cpu_arm(images_per_core=1, cfg_id=args.config)
        t0 = time.perf_counter()
        for _ in range(steps):
            base, _wall = cpu_arm(images_per_core=1, cfg_id=args.config)
            vals.append(base["value"])
        dt = time.perf_counter() - t0
        v = float(np.median(vals))
        base["value"] = v
        print(json.dumps({
            "impl": "reference", "metric": cfg["metric"], "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "note": "each step = one image per host core through the reference's "
                       "get_full_report_data (CPU only; the reference has no GPU path and no OpenMP)"},
            "cpu_baseline": base, "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 meanwhile (NCCL's version banner, the
    # reference's printf in the cpu_baseline leg) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from photohive_dsp_b200.batch import Context, MultiContext, flat_layout, make_params
    from tools.synth import Generator

    if not torch.cuda.is_available():
        os.write(real_stdout, (json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}) + "\n").encode())
        return 2
    n_local = args.gpus if in_process else 1           # GPUs this process drives
    if in_process and torch.cuda.device_count() < args.gpus:
        os.write(real_stdout, (json.dumps({"error": f"--gpus {args.gpus} but only {torch.cuda.device_count()} visible"}) + "\n").encode())
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on (and first-touch the pinned staging buffers from) the CPUs next to this GPU
    numa = "not set (one process drives every GPU)"
    if not in_process:
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "gpu-local cpus"
        except Exception as e:  # affinity is a host-side nicety, never a reason to fail the measurement
            numa = f"not set ({type(e).__name__})"
    if world > 1:
        # keep stdout to the one JSON line: NCCL's version banner goes there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        for d in range(n_local):
            torch.cuda.synchronize(local_rank + d if in_process else local_rank)

    params = make_params(**cfg["params"])
    nbox = cfg["nbox"]
    lay = flat_layout(params, nbox)
    stride = w * h * 3
    box_area = sum((b - t) * (r - l) for t, b, l, r in boxes_for(w, h, nbox))
    boxes_np = np.array([boxes_for(w, h, nbox)] * (batch * n_local), np.int32) if nbox else None
    boxes_ptr = boxes_np.ctypes.data if nbox else None

    # ---- device-resident shards: one per GPU this process drives ------------------------------------------------
    shards = []
    for d in range(n_local):
        di = local_rank + d
        ddev = torch.device("cuda", di)
        shard_index = rank if not in_process else d
        with torch.cuda.device(ddev):
            gen = Generator(w, h, ddev)
            first = FIRST_SEED + shard_index * batch   # every shard gets its own slice of the seed sequence
            images = gen.batch(batch, first)           # [B,H,W,3] uint8 in HBM, far larger than L2
            del gen
            torch.cuda.empty_cache()
            records = torch.empty((batch, lay.record_bytes), dtype=torch.uint8, device=ddev)
        shards.append(dict(ctx=Context(di), images=images, records=records, dev=ddev))
    ctx, images, records = shards[0]["ctx"], shards[0]["images"], shards[0]["records"]

    def shard_step(sh):
        sh["ctx"].get_reports_raw(sh["images"].data_ptr(), batch, w, h, stride, params, sh["records"].data_ptr(),
                                  boxes_ptr=boxes_ptr, max_boxes=nbox)

    def step():
        if n_local == 1:
            shard_step(shards[0])
            return
        ts = [threading.Thread(target=shard_step, args=(sh,)) for sh in shards[1:]]  # ctypes calls release the GIL
        for t in ts:
            t.start()
        shard_step(shards[0])
        for t in ts:
            t.join()

    for _ in range(max(args.warmup, 3)):
        step()
    # outside the timed region: record 0 of the default workload is the committed golden case (unmodified reference)
    checked = None
    if rank == 0 and args.config == 3:
        if os.environ.get("PHD_BENCH_SKIP_CHECK"):   # measurement-only experiment builds (tools/ab_run.sh)
            checked = "SKIPPED (PHD_BENCH_SKIP_CHECK set: not a valid bench line)"
        else:
            checked = check_record_zero(records[0].cpu().numpy(), lay)
    clk_path = os.path.join(tempfile.gettempdir(), f"phd_clocks_{rank}.csv")
    sampler = clocks_sampler_start(clk_path) if rank == 0 else None
    barrier()
    stage = {}
    stage_launches = {}
    launches = 0
    dev_ms = 0.0
    fused = False
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        step_ms = 0.0
        for i, sh in enumerate(shards):
            ms, nl = sh["ctx"].last_timing()
            launches += nl
            step_ms = max(step_ms, ms["total"])        # the GPUs of one process run concurrently: the slowest counts
            if i == 0:
                for k, v in ms.items():
                    stage[k] = stage.get(k, 0.0) + v
                for k, v in sh["ctx"].last_stage_launches().items():
                    stage_launches[k] = stage_launches.get(k, 0) + v
                fused = fused or sh["ctx"].last_fused()
        dev_ms += step_ms
    barrier()
    wall = time.perf_counter() - t0
    # CUDA-event time of the K steps on the pipeline's own stream; max over ranks
    t = torch.tensor([dev_ms / 1000.0, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])

    # ---- end to end: pinned host input -> H2D -> pipeline -> D2H records -> gathered on the host --------------
    n_shards = world * n_local
    host_in = torch.empty((eb * n_local, h, w, 3), dtype=torch.uint8).pin_memory()
    for d, sh in enumerate(shards):
        host_in[d * eb:(d + 1) * eb].copy_(sh["images"][:eb])
    eboxes = np.array([boxes_for(w, h, nbox)] * (eb * n_local), np.int32) if nbox else None
    eboxes_ptr = eboxes.ctypes.data if nbox else None
    shm_path, gather = None, "single GPU: the call's own D2H copy"
    if world > 1:
        # every rank's device writes its records into its slice of ONE pinned host array: gather = closing barrier
        name = f"phd_bench_{os.environ.get('MASTER_PORT', '0')}_records"
        if rank == 0:
            all_out, shm_path, pinned = shared_pinned_array(name, world * eb * lay.record_bytes, create=True)
        dist.barrier()
        if rank != 0:
            all_out, shm_path, pinned = shared_pinned_array(name, world * eb * lay.record_bytes, create=False)
        all_out = all_out.view(world, eb, lay.record_bytes)
        host_out = all_out[rank]
        gather = ("every rank's records land in its slice of one shared pinned host array (POSIX shm, "
                  f"cudaHostRegister {'ok' if pinned else 'FAILED: pageable'}); rank 0 holds all after the closing barrier")
        multi = None
    elif in_process:
        host_out = torch.empty((eb * n_local, lay.record_bytes), dtype=torch.uint8).pin_memory()
        multi = MultiContext(list(range(local_rank, local_rank + n_local)))
        gather = "phd_get_reports_u8_multi: one host thread per GPU, each device writes its range of one pinned host array"
    else:
        host_out = torch.empty((eb, lay.record_bytes), dtype=torch.uint8).pin_memory()
        multi = None

    def e2e_step():
        if multi is not None:
            multi.get_reports_raw(host_in.data_ptr(), eb * n_local, w, h, stride, params, host_out.data_ptr(),
                                  boxes_ptr=eboxes_ptr, max_boxes=nbox)
        else:
            ctx.get_reports_raw(host_in.data_ptr(), eb, w, h, stride, params, host_out.data_ptr(),
                                boxes_ptr=eboxes_ptr, max_boxes=nbox)

    for _ in range(3):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()   # closes the gather: every rank's last records are in the shared array
    e2e_wall = time.perf_counter() - t1
    te = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    if multi is not None:
        same = all(bool(torch.equal(host_out[d * eb:(d + 1) * eb].to(sh["dev"]), sh["records"][:eb]))
                   for d, sh in enumerate(shards))
    else:
        same = bool(torch.equal(host_out.to(dev), records[:eb]))
    gathered_ok = None
    if world > 1:
        flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(int(flag[0]))
        if rank == 0:   # rank 0 reads every rank's slice: all records present and valid (status field 0, N > 0)
            from photohive_dsp_b200.batch import view_records
            v = view_records(all_out.view(world * eb, lay.record_bytes).numpy(), lay)
            gathered_ok = bool((v.palette_n > 0).all())
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(all_out.data_ptr())
        if rank == 0 and shm_path:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
    if multi is not None:
        multi.close()
    if sampler is not None:
        sampler.terminate()
        sampler.wait()

    if rank == 0:
        peak, peak_src = measured_peak()
        n_img_job = batch * n_shards * args.steps
        value = n_img_job / dev_s
        kb = kernel_bytes(w, h, box_area, fused)
        per_kernel = {k: stage.get(k, 0.0) for k in kb}
        dom = max(per_kernel, key=per_kernel.get)
        dom_s = per_kernel[dom] / 1000.0
        n_img_rank = batch * args.steps
        achieved = kb[dom] * n_img_rank / dom_s / 1e9
        n_dom_launches = max(stage_launches.get(dom, 1), 1)
        traffic = None  # dram read+write bytes per launch of that kernel from the committed ncu capture, if any
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f).get(dom)
            if tj and args.config == 3 and not fused:  # bytes per 1080p image in the capture -> bytes per launch of this run
                traffic = tj["dram_bytes_per_image"] * n_img_rank / n_dom_launches
        except Exception:
            pass
        ab = algo_bytes(w, h, box_area)
        pipe_achieved = ab * n_img_rank / (dev_ms / 1000.0) / 1e9
        line = {
            "metric": cfg["metric"], "value": value, "unit": "images/s", "n_gpus": n_shards, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 pixels, int32/int64 fixed-point sums, f32 FFT",
            "data": "synthetic",
            "config": {"workload": workload, "baseline_config": args.config,
                       "l2": f"inputs larger than L2 ({batch * stride / 1e9:.1f} GB per GPU per step)",
                       "timing": "CUDA events on the library's stream around each step, summed over the K steps, max over ranks",
                       "wall_ms_per_step": 1000.0 * wall_s / args.steps, "sub_batch": os.environ.get("PHD_SUB_BATCH", "auto"),
                       "launch": "one process drives every GPU (phd_get_reports_u8_multi)" if in_process else
                                 ("one process per GPU (torchrun)" if world > 1 else "one process, one GPU"),
                       "front_rows_fused": fused, "record0_check": checked},
            "e2e": {"value": eb * n_shards * args.steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": eb * stride * n_local,
                    "d2h_bytes_per_step": eb * lay.record_bytes * n_local, "batch": eb, "records_identical_to_device_run": same,
                    "path": "phd_get_reports_u8_multi (C ABI)" if in_process else "phd_get_reports_u8 (C ABI) with pinned host buffers",
                    "gather": gather, "gathered_records_valid": gathered_ok, "cpu_affinity": numa},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom + (" (front end + row FFT, one launch)" if fused and dom == "frontend" else ""),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_image": kb[dom], "launches": n_dom_launches,
                         "algorithmic_bytes_per_launch": kb[dom] * n_img_rank / n_dom_launches,
                         "avg_launch_ms": per_kernel[dom] / n_dom_launches,
                         "note": "bound by the LSU data pipe (ncu: 78 % of peak), not by HBM: see DESIGN.md section 6"},
            "roofline_pipeline": {"algo_bytes_per_image": ab, "achieved": pipe_achieved, "peak": peak,
                                  "unit": "GB/s", "frac": pipe_achieved / peak},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "clocks": clocks_summary(clk_path, local_rank),
        }
        if n_shards == 1 and not args.no_cpu:
            try:
                # ~20 s of CPU work on a 16-core host at 1080p; one image per core for the big shapes
                line["cpu_baseline"], _ = cpu_arm(images_per_core=3 if w * h <= 2_100_000 else 1, cfg_id=args.config)
            except Exception as e:  # the baseline is a reported figure; never let it hide the measurement
                line["cpu_baseline"] = {"error": str(e)[:200]}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
