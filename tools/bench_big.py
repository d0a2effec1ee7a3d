"""Device-resident throughput of the two large BASELINE configs only (A/B helper): python tools/bench_big.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv.append("--only-big")
exec(open(os.path.join(ROOT, "tools", "bench_configs.py")).read())
