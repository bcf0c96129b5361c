#!/bin/bash
# usage: scripts/ncu_case.sh <bringup-case> <kernel-regex> <out-name>   (run under gpurun, one GPU)
set -u
CASE=$1; KREGEX=$2; OUT=$3
python scripts/gpu_bringup.py "$CASE" > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 2 -c 2 -f -o gpurun_out/${OUT} \
    python scripts/gpu_bringup.py "$CASE" > gpurun_out/${OUT}_ncu.log 2>&1
tail -3 gpurun_out/${OUT}_plain.log
tail -3 gpurun_out/${OUT}_ncu.log
