"""Turns the gpurun_out/ captures of tools/profile_round.sh into the committed summaries under profiles/.
usage (here, no GPU needed): python tools/profile_summarize.py TAG [images in the --set full capture, default 32]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
nimg_full = int(sys.argv[2]) if len(sys.argv) > 2 else 32
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")


def short(name):
    for k in ("k_pixels", "k_palette_select", "k_palette_ties", "k_rows_t", "k_rows_generic", "k_cols_t", "k_cols_generic",
              "k_finalize", "k_sharpness", "k_rgb_stats"):
        if k in name:
            return k
    return name[:40]


lines = []
# ---- launch list --------------------------------------------------------------------------------
lp = os.path.join(go, f"{tag}_launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    per = {}
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("kernel,grid,block,gpu__time_duration_us\n")
        for r in rows[1:]:
            t = float(r[vi].replace(",", ""))
            t_us = t / 1000.0 if r[ui] in ("ns", "nsecond") else (t if r[ui] in ("us", "usecond") else t * 1000.0)
            k = short(r[ki])
            f.write(f"{k},\"{r[gi]}\",\"{r[bi]}\",{t_us:.3f}\n")
            a = per.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += t_us
    tot = sum(v[1] for v in per.values())
    lines += [f"## Launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --batch 256 --steps 2 --warmup 3`",
              "", "Cold-cache, serialised times: compare SHARES, not absolutes.", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {n} | {t:.0f} | {100 * t / tot:.1f} % |")
    lines.append("")

# ---- full capture -------------------------------------------------------------------------------
rp = os.path.join(go, f"{tag}_full.ncu-rep")
traffic = {}
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with open(os.path.join(go, f"{tag}_full_raw.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("smsp__inst_executed.sum", "warp instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data pipe % of peak"),
            ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
    lines += [f"## `ncu --set full --clock-control none` ({tag}): one launch each, {nimg_full}-image call (`tools/prof_driver.py {nimg_full}`)", ""]
    names = [short(r[idx["Kernel Name"]]) for r in rows[2:]]
    lines.append("| metric | " + " | ".join(names) + " |")
    lines.append("|---|" + "---|" * len(names))
    for m, label in want:
        if m in idx:
            lines.append(f"| {label} ({units[idx[m]]}) | " + " | ".join(r[idx[m]] for r in rows[2:]) + " |")
    lines.append("")

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    stage_of = {"k_pixels": "frontend", "k_rows_t": "fft_rows", "k_cols_t": "fft_cols_blur"}
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        if k in stage_of:
            # the image index is the grid's y dimension in all three kernels
            gtxt = r[idx["Grid Size"]] if "Grid Size" in idx else ""
            dims = [int(x) for x in gtxt.strip("()").replace(" ", "").split(",") if x]
            n_img = dims[1] if len(dims) > 1 else nimg_full
            grid = gtxt
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            traffic[stage_of[k]] = {"kernel": k, "grid": grid, "images_in_launch": n_img,
                                    "dram_bytes_per_launch": rd + wr, "dram_bytes_per_image": (rd + wr) / n_img}
    json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)

with open(os.path.join(out_dir, f"{tag}_summary.md"), "w") as f:
    f.write(f"# Profile summary {tag}\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
