#!/usr/bin/env bash
# usage: tools/ncu_one.sh <tag> <kernel-regex> <skip> [script args...]  -> gpurun_out/<tag>.ncu-rep
TAG="$1"; RE="$2"; SKIP="$3"; shift 3
mkdir -p gpurun_out
MEDGP_STREAMS=1 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s "$SKIP" -c 1 \
    -f -o "gpurun_out/$TAG" python tools/profile_step.py "$@" > "gpurun_out/$TAG.log" 2>&1
ls -la "gpurun_out/$TAG.ncu-rep"
