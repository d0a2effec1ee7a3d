"""Prints the key fields of a bench.py JSON line: python tools/bench_line.py file.json [label]"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
st = d.get("stage_ms_per_step", {})
print(sys.argv[2] if len(sys.argv) > 2 else "", f"{d['value']:.0f} img/s  e2e {d['e2e']['value']:.0f}  ",
      "  ".join(f"{k}={v:.3f}" for k, v in st.items()))
