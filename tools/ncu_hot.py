"""Top stall-sample SASS instructions of one kernel from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME | python tools/ncu_hot.py [N]"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None
data = []
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
iwfi = hdr.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hdr else None
tot = sum(int(r[isamp]) for r in data)
base = int(data[0][ia], 16)
print("total samples", tot, "instructions", len(data), "executed", sum(int(r[iexec]) for r in data))
cum = 0
for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
    cum += int(r[isamp])
    extra = f" wf {r[iwf]}/{r[iwfi]}" if iwf is not None and r[iwf] not in ("0", "") else ""
    print(f"{int(r[ia],16)-base:6x} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}% cum {100*cum/tot:5.1f}%  {r[isrc].strip()[:90]}{extra}")
