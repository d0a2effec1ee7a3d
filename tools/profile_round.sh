#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the ncu passes whose summaries are committed under profiles/.
#   tools/profile_round.sh TAG launches|full        (one ncu pass per gpurun call)
# launches: every launch of our kernels in a (reduced-batch) bench.py run with its device time -> gpurun_out/TAG_launches.csv
# full:     one `--set full` capture of the three streaming kernels on a 32-image call        -> gpurun_out/TAG_full.ncu-rep
# Each ncu pass only runs after the same command exited 0 without ncu (B200_PROFILING.md).
TAG=${1:-r1}
WHAT=${2:-launches}
K='k_pixels|k_palette|k_rows|k_cols|k_finalize|k_sharpness|k_rgb_stats'
BENCH="python bench.py --batch 256 --steps 2 --warmup 3 --no-cpu"
if [ "$WHAT" = launches ]; then
    $BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err &&
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 700 --csv \
        --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_bench_ncu.log 2>&1
    tail -2 gpurun_out/${TAG}_bench_ncu.log
else
    python tools/prof_driver.py 32 > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'k_pixels|k_rows|k_cols' -c 3 \
        -o gpurun_out/${TAG}_full -f python tools/prof_driver.py 32 > gpurun_out/${TAG}_full_ncu.log 2>&1
    tail -2 gpurun_out/${TAG}_full_ncu.log
fi
