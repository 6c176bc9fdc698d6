#!/bin/bash
# ncu --set full capture (with source) of one kernel of one op: tools/gpu_prof_op.sh <tag> <kernel-regex> <op> <shape>
TAG=$1; KRE=$2; OP=$3; SHAPE=$4
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$KRE --launch-skip 2 --launch-count 1 \
  -f -o gpurun_out/${TAG} python tools/run_op.py --op $OP --shape $SHAPE --iters 4 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu $TAG exit $?"; tail -2 gpurun_out/${TAG}_ncu.log
