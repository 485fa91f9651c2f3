#!/bin/bash
# Multi-GPU measurements of one round on an N-GPU box:  gpurun --gpus N -- 'bash tools/run_multi_gpu.sh rNN N'
#   <tag>_pcie_nN.txt      host-link ceiling with 1, 2, 4, .. N GPUs active at once (tools/pcie_probe_multi.py)
#   <tag>_slab_tests_nN.log  the multi-rank slab tests over NCCL / NVLink (tests/test_gpu_slab.py)
#   <tag>_cfg3_nN.json     bench.py --gpus N (view sharding, weak scaling; every e2e mode and transport)
#   <tag>_cfg5_nK.json     bench.py --workload cfg5 --gpus K for K = N (and the smaller K given as extra arguments)
set -u
tag=${1:-rXX}; n=${2:-8}; shift 2
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
nvidia-smi -L > $out/${tag}_gpus_n$n.txt 2>&1; nproc >> $out/${tag}_gpus_n$n.txt; free -g | head -2 >> $out/${tag}_gpus_n$n.txt
nvidia-smi topo -m >> $out/${tag}_gpus_n$n.txt 2>&1
if [ -z "${SLAB_ONLY:-}" ]; then
run $n tools/pcie_probe_multi.py > $out/${tag}_pcie_n$n.txt 2> $out/${tag}_pcie_n$n.err
python -m pytest tests/test_gpu_slab.py -m gpu -x -q > $out/${tag}_slab_tests_n$n.log 2>&1; tail -3 $out/${tag}_slab_tests_n$n.log
run $n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu > $out/${tag}_cfg3_n$n.json 2> $out/${tag}_cfg3_n$n.err; tail -c 300 $out/${tag}_cfg3_n$n.err
fi
for k in $n "$@"; do
    # NVLink byte counters of GPU 0 around the whole run (warm-up 2 + 3 timed convolutions): nvidia-smi, when NVML's field values are refused
    nvidia-smi nvlink -gt d -i 0 > $out/${tag}_nvlink_before_n$k.txt 2>&1
    run $k bench.py --workload cfg5 --gpus $k --steps 3 --warmup 2 > $out/${tag}_cfg5_p2p_n$k.json 2> $out/${tag}_cfg5_p2p_n$k.err; tail -c 300 $out/${tag}_cfg5_p2p_n$k.err
    nvidia-smi nvlink -gt d -i 0 > $out/${tag}_nvlink_after_n$k.txt 2>&1
    MVSIM_SLAB_NCCL=1 run $k bench.py --workload cfg5 --gpus $k --steps 3 --warmup 2 > $out/${tag}_cfg5_nccl_n$k.json 2> $out/${tag}_cfg5_nccl_n$k.err
done
ls -la $out | tail -20
