#!/bin/bash
# Offline (no GPU) size of the front end's hot loop: SASS instructions between the first and the last shared-memory
# atomic of the unrolled 16-pixel block of k_pixels<256, false, ...>, plus registers / shared memory from ptxas.
# The kernel is bound by instruction issue, so this number tracks its run time (one instruction per pixel is about
# 0.7 % of the kernel); ptxas is moody about the block (a few live registers more once made it 9 % longer), so check it
# after every change to frontend.cu / pixel_cells.cuh.   usage: tools/sass_span.sh [extra nvcc flags]
cd "$(dirname "$0")/../photohive_dsp_b200/csrc" || exit 1
OUT=${TMPDIR:-/tmp}/phd_fe_span
nvcc -ccbin /usr/bin/g++ -I../../include -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v "$@" \
    -c frontend.cu -o $OUT.o 2>&1 | grep -A2 "k_pixelsILi256ELb0" | grep -E "Used" | sed 's/ptxas info    : //'
cuobjdump -sass $OUT.o | awk '/Function :/{name=$3} name ~ /k_pixelsILi256ELb0/ && /^ +\/\*[0-9a-f]{4}\*\//{print}' > $OUT.sass
first=$(grep -n ATOMS $OUT.sass | head -1 | cut -d: -f1)
last=$(grep -n ATOMS $OUT.sass | sed -n 48p | cut -d: -f1)
echo "hot block: $((last - first)) instructions between the first and the 48th ATOMS (15 pixels); kernel: $(wc -l < $OUT.sass)"
