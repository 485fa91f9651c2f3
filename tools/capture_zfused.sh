#!/bin/bash
# A/B capture of the fused z pass variants (to be run on the GPU box): timings first, then one --set full report per variant.
#   gpurun -- 'bash tools/capture_zfused.sh rNN'
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
MVSIM_Z_KERNEL=3 python tools/time_view.py 12 > $out/${tag}_time.txt 2>&1
MVSIM_Z_KERNEL=1 python tools/time_view.py 12 >> $out/${tag}_time.txt 2>&1
cat $out/${tag}_time.txt
: > $out/${tag}_zkernels.md
MVSIM_Z_KERNEL=3 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ZFused" -c 1 -f -o $out/${tag}_zpoly \
    python tools/prof_conv.py 1 > $out/${tag}_ncu_zpoly.log 2>&1
python tools/summarize_ncu.py kernels $out/${tag}_zpoly.ncu-rep >> $out/${tag}_zkernels.md 2>&1
MVSIM_Z_KERNEL=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ZFused" -c 1 -f -o $out/${tag}_zdec \
    python tools/prof_conv.py 1 > $out/${tag}_ncu_zdec.log 2>&1
python tools/summarize_ncu.py kernels $out/${tag}_zdec.ncu-rep >> $out/${tag}_zkernels.md 2>&1
cat $out/${tag}_zkernels.md
