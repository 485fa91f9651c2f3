#!/bin/bash
# A/B timings of one knob on the GPU box: bash tools/time_ab.sh <tag> <ENVVAR> <value_a> <value_b> [reps] [workload]
set -u
tag=$1; var=$2; a=$3; b=$4; reps=${5:-12}; wl=${6:-cfg3}
out=gpurun_out; mkdir -p $out
env $var=$a python tools/time_view.py $reps $wl > $out/${tag}_ab.txt 2>&1
env $var=$b python tools/time_view.py $reps $wl >> $out/${tag}_ab.txt 2>&1
cat $out/${tag}_ab.txt
