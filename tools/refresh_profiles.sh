#!/usr/bin/env bash
# Everything profiles/ is made of, in one gpurun call (1 GPU):
#   tools/refresh_profiles.sh <tag>        -> gpurun_out/<tag>_*
# then, back in the container:  python tools/refresh_profiles_post.py <tag>
set -u
TAG="${1:-r01}"
mkdir -p gpurun_out
# 1. launch list of two C2 steps (durations), single stream
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file "gpurun_out/${TAG}_launches_c2_step.csv" python tools/profile_step.py 2 > /dev/null 2>&1
# 2. DRAM bytes of every launch of one C2 step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file "gpurun_out/${TAG}_dram_traffic_c2_step.csv" python tools/profile_step.py 1 > /dev/null 2>&1
# 3. one --set full capture per hot kernel
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 \
      -f -o "gpurun_out/${TAG}_ncu_$1" python tools/profile_step.py 1 > "gpurun_out/${TAG}_ncu_$1.log" 2>&1
}
cap grad 'k_grad$' 0
cap lauum 'k_lauum' 0
cap assemble 'k_assemble' 0
cap diag4 'k_potrf_diag' 4
cap panel3 'k_potrf_panel' 3
cap trtri5 'k_trtri_row' 4
# 4. the bench lines themselves (never under a profiler)
python bench.py --impl reference > "gpurun_out/${TAG}_bench_reference_arm.json" 2> /dev/null
python bench.py > "gpurun_out/${TAG}_bench_n1.json" 2> "gpurun_out/${TAG}_bench_n1.err"
python tools/bench_predict.py 16 500 > "gpurun_out/${TAG}_online_imputation.json" 2> /dev/null
python tools/bench_cholesky.py > "gpurun_out/${TAG}_cholesky_n4000.json" 2> /dev/null
python tools/bench_latency.py 2> /dev/null | tail -1 > "gpurun_out/${TAG}_latency.json"
python tools/bench_c3.py 256 5 3 2> /dev/null | tail -1 > "gpurun_out/${TAG}_c3_ragged.json"
ls -la gpurun_out | tail -30
