#!/bin/bash
# Runs ON THE GPU BOX: bench stage times of several library variants back to back (ab_libs/*.so, see tools/ab_build.sh).
#   tools/ab_run.sh "name1 name2 ..." "3 2"     (variant names, BASELINE configs)
for v in $1; do
  for c in $2; do
    b=2048; [ $c = 2 ] && b=256; [ $c = 4 ] && b=96
    PHD_LIB_PATH=$PWD/ab_libs/$v.so python bench.py --no-cpu --steps 3 --config $c --batch $b --e2e-batch 8 > gpurun_out/ab_${v}_c$c.json 2> gpurun_out/ab_${v}_c$c.err || tail -3 gpurun_out/ab_${v}_c$c.err
    python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_${v}_c$c.json"))
    s = d["stage_ms_per_step"]
    print("$v c$c: %.0f img/s  fe %.2f rows %.2f cols %.2f ties %.2f sel %.2f fin %.2f  same=%s" % (d["value"], s["frontend"], s["fft_rows"], s["fft_cols_blur"], s["palette_ties"], s["palette_select"], s["finalize"], d["e2e"]["records_identical_to_device_run"]))
except Exception as e:
    print("$v c$c failed", e)
P
  done
done
