#!/bin/bash
# per-launch duration, tensor-pipe instruction count and tensor-pipe activity of one reduced C3 step
# usage: tools/ncu_dmma_list.sh out_prefix lib...
out=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  MEDGP_LIB=$lib ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size \
     --clock-control none --csv --log-file ${out}_${name}_dmma.csv python tools/profile_c3.py 1 > /dev/null 2>&1
done
