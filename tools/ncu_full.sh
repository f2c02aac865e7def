#!/usr/bin/env bash
# ncu --set full captures of the hot kernels of one bench step (run under gpurun, 1 GPU).
# usage: tools/ncu_full.sh <tag>   -> gpurun_out/<tag>_<kernel>.ncu-rep
set -u
TAG="${1:-r1}"
mkdir -p gpurun_out
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 \
      -f -o "gpurun_out/${TAG}_$1" python tools/profile_step.py 1 > "gpurun_out/${TAG}_$1.log" 2>&1
}
cap grad 'k_grad$' 0
cap lauum 'k_lauum' 0
cap assemble 'k_assemble' 0
cap diag 'k_potrf_diag' 4
cap panel 'k_potrf_panel' 3
cap trtri 'k_trtri_row' 6
ls -la gpurun_out/*.ncu-rep
