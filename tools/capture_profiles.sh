#!/bin/bash
# Profile capture for one round, to be run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh rNN').
# One --set full report PER KERNEL of a config-3 view (tools/prof_conv.py), each 10-15 MB, so that the directory stays below the
# runner's 64 MiB limit (a single report of all nine kernels is ~65 MB: the round-1 capture of the final code was lost that way).
# Text summaries are written first; the reports themselves are kept in priority order while they fit.
#   gpurun_out/<tag>_launches.md   launch list of `bench.py --steps 1 --warmup 1 --no-cpu` (gpu__time_duration.sum)
#   gpurun_out/<tag>_kernels.md    per-kernel summary (tools/summarize_ncu.py)
#   gpurun_out/<tag>_<kernel>.ncu-rep   read here with  ncu -i ... --page source --csv --print-source sass
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python tools/prof_conv.py 1 > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; cat $out/${tag}_plain.log; exit 1; }
# name : regex on the demangled kernel name : launches of that name to skip (the PSF's x / y passes come before the image's)
kernels="zfused:ZFused:0 sample:extract_:0 rotate:rotate_attenuate_kernel:0 xfwd:XFwd:1 yfwd:StridedFwd:1 xinv:XInv:0 yinv:StridedInv:0"
: > $out/${tag}_kernels.md
for k in $kernels; do
    name=${k%%:*}; rest=${k#*:}; rx=${rest%%:*}; skip=${rest##*:}
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 -f -o $out/${tag}_$name \
        python tools/prof_conv.py 1 > $out/${tag}_ncu_$name.log 2>&1
    python tools/summarize_ncu.py kernels $out/${tag}_$name.ncu-rep >> $out/${tag}_kernels.md 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_launch.log 2>&1
python tools/summarize_ncu.py launches $out/${tag}_launches.csv > $out/${tag}_launches.md 2>&1
# keep the directory under the limit: drop reports from the END of the priority list first
for k in $(echo $kernels | tr ' ' '\n' | tac); do
    [ "$(du -sm $out | cut -f1)" -lt 58 ] && break
    name=${k%%:*}
    echo "dropping ${tag}_$name.ncu-rep to stay below 64 MiB" >> $out/${tag}_capture.log
    rm -f $out/${tag}_$name.ncu-rep
done
du -sh $out
