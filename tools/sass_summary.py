"""Offline (no GPU) evidence of what the built library contains: per kernel family, the count of the SASS mnemonics
that identify the Blackwell-era paths (B200_PROFILING.md, "What proves a Blackwell-native kernel") and the classic
ones.  usage: python tools/sass_summary.py [path to libreport_data.so] > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "photohive_dsp_b200", "PhotoHive_DSP_lib", "libreport_data.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "ATOMS", "REDG", "ATOMG", "RED", "IDP", "VIMNMX3",
         "FMNMX3", "MUFU", "BAR", "LDGSTS", "LDL", "STL", "HMMA", "UTCHMMA", "LDTM"]
FAMILIES = ["k_pixels", "k_front_rows", "k_palette_select", "k_palette_ties", "k_rows_t", "k_rows_generic", "k_cols_t",
            "k_cols_generic", "k_sharpness", "k_finalize", "k_bin_map", "k_rgb_stats", "k_ingest", "k_build", "k_group_sweep",
            "k_twiddles", "k_pass_twiddles", "k_bluestein", "k_pixels_f64", "k_palette_f64"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
fam = None
counts = collections.defaultdict(collections.Counter)
ninst = collections.Counter()
nfun = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        fam = next((f for f in sorted(FAMILIES, key=len, reverse=True) if f in name), "other")
        nfun[fam] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fam:
        op = m.group(1)
        ninst[fam] += 1
        for wname in WATCH:
            if op == wname:
                counts[fam][wname] += 1
cols = [w for w in WATCH if any(counts[f][w] for f in counts)]
print(f"# SASS summary of `{os.path.relpath(lib, ROOT)}`\n")
print(f"`cuobjdump -sass`, architectures in the fatbin: {', '.join(arch)}.  One row per kernel family (all template "
      "instantiations summed); columns are instruction counts in the machine code, not executed counts.\n")
print("`UBLKCP` = cp.async.bulk (the TMA copy engine, 1-D), `SYNCS` = mbarrier operations, `FFMA2/FADD2/FMUL2` = packed FP32x2 "
      "(sm_100), `ATOMS` = shared-memory atomics, `REDG/ATOMG` = global reductions, `IDP` = dp2a/dp4a, `LDL/STL` = local-memory "
      "(spill) accesses.  `UTMASTG` = cp.async.bulk.tensor store (tensor-map TMA: the row kernel's transposed output); no `UTMALDG`, `HMMA`/`UTC*MMA` (tensor cores) -- nothing here is a dense contraction.\n")
print("| kernel family | instantiations | instructions | " + " | ".join(cols) + " |")
print("|---|---|---|" + "---|" * len(cols))
for f in sorted(ninst, key=lambda k: -ninst[k]):
    print(f"| {f} | {nfun[f]} | {ninst[f]} | " + " | ".join(str(counts[f][c]) for c in cols) + " |")
