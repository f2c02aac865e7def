#!/bin/bash
# ncu --set full captures of the dataflow kernels (one n = 4000 matrix in flight), summarised on the box
for k in potrf trtri; do
  MEDGP_GRAPHS=0 ncu --set full --clock-control none --import-source on -k "regex:k_${k}_flow" -s 0 -c 1 -f \
      -o gpurun_out/r02_ncu_flow_${k}_n4000 python tools/longstay_one.py 1 4000 2 1 > gpurun_out/r02_ncu_flow_${k}_n4000.log 2>&1
  rep=gpurun_out/r02_ncu_flow_${k}_n4000.ncu-rep
  if [ -f $rep ]; then
    (python tools/ncu_summary.py $rep; echo; echo "-- hottest SASS lines (tools/ncu_hot.py) --"; python tools/ncu_hot.py $rep 2.0; echo; echo "-- executed instruction mix (tools/ncu_instmix.py) --"; python tools/ncu_instmix.py $rep | head -14) > gpurun_out/r02_ncu_flow_${k}_n4000.summary.txt
    rm -f $rep
  fi
  tail -3 gpurun_out/r02_ncu_flow_${k}_n4000.log
done
