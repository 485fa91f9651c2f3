#!/bin/bash
# Profile capture for one round, to be run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh rNN').
# Writes small TEXT summaries into gpurun_out/ first and keeps the .ncu-rep files only while the directory stays
# below the runner's 64 MiB limit (a single --set full report of the nine kernels of a view is ~65 MB: the round-1
# capture of the final code was lost that way).
#   gpurun_out/<tag>_launches.md   launch list of `bench.py --steps 1 --warmup 1 --no-cpu` (gpu__time_duration.sum)
#   gpurun_out/<tag>_kernels.md    per-kernel summary of one view (tools/prof_conv.py), two reports of <= 5 kernels each
#   gpurun_out/<tag>_z_source.csv  per-instruction counts / stall samples of the fused z pass (source page)
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python tools/prof_conv.py 1 > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
K='regex:fft_kernel|extract_kernel|rotate_attenuate'
ncu --set full --clock-control none --import-source on -k "$K" -c 5 -f -o $out/${tag}_a python tools/prof_conv.py 1 > $out/${tag}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 5 -c 4 -f -o $out/${tag}_b python tools/prof_conv.py 1 > $out/${tag}_ncu_b.log 2>&1
{ python tools/summarize_ncu.py kernels $out/${tag}_a.ncu-rep; python tools/summarize_ncu.py kernels $out/${tag}_b.ncu-rep; } > $out/${tag}_kernels.md 2>&1
ncu -i $out/${tag}_b.ncu-rep --page source --csv --print-source sass -k regex:ZFused 2>/dev/null | head -7000 > $out/${tag}_z_source.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_launch.log 2>&1
python tools/summarize_ncu.py launches $out/${tag}_launches.csv > $out/${tag}_launches.md 2>&1
# keep the directory under the limit: drop the largest reports first
while [ "$(du -sm $out | cut -f1)" -ge 60 ]; do
    big=$(ls -S $out/*.ncu-rep 2>/dev/null | head -1)
    [ -z "$big" ] && break
    echo "dropping $big to stay below 64 MiB" >> $out/${tag}_capture.log
    rm -f "$big"
done
du -sh $out
