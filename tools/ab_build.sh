#!/bin/bash
# Builds a variant of libreport_data.so for A/B runs on one box: tools/ab_build.sh NAME "-DPHD_KNOB=1 ..." -> ab_libs/NAME.so
# (pick it at run time with PHD_LIB_PATH=ab_libs/NAME.so; ab_libs/ is git-ignored but travels with gpurun).
# The in-tree library is restored to the default build afterwards: leaving the variant in place once made every later
# measurement of "HEAD" a measurement of the variant.
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_libs
PHD_NVCC_EXTRA="$2" python -m photohive_dsp_b200.build --force > /dev/null
cp photohive_dsp_b200/PhotoHive_DSP_lib/libreport_data.so ab_libs/$1.so
echo "built ab_libs/$1.so with [$2]"
if [ -n "$2" ]; then
    python -m photohive_dsp_b200.build --force > /dev/null
    echo "restored the default build in-tree"
fi
