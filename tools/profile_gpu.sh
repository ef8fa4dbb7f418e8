#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  $1 = tag, $2 = backend (0 simt, 1 tcgen05),
# $3 = kernel regex for the full capture, $4 = matching launches to skip, $5 = launches to capture.
TAG=${1:-r01}; BACKEND=${2:-0}; KRE=${3:-conv_simt_kernel}; SKIP=${4:-121}; CNT=${5:-3}
mkdir -p gpurun_out
export SB200_BACKEND=$BACKEND
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c ${CNT} \
    -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
