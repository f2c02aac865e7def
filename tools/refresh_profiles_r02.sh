#!/usr/bin/env bash
# Everything profiles/r02_* is made of, in one gpurun call (1 GPU):
#   tools/refresh_profiles_r02.sh [tag]      -> gpurun_out/<tag>_*
# then, back in the container:  python tools/refresh_profiles_post_r02.py [tag]
set -u
TAG="${1:-r02}"
mkdir -p gpurun_out
# 1. launch list (durations) of two steps of the reduced C3 cohort, single stream
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file "gpurun_out/${TAG}_launches_c3_step.csv" python tools/profile_c3.py 2 > "gpurun_out/${TAG}_profile_c3.json" 2> /dev/null
# 2. DRAM bytes of every launch of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file "gpurun_out/${TAG}_dram_traffic_c3_step.csv" python tools/profile_c3.py 1 > /dev/null 2>&1
# 3. one --set full capture per hot kernel (C3 cohort; mid-factorisation launches)
cap() {  # name regex skip script...
  local name="$1" re="$2" skip="$3"; shift 3
  MEDGP_STREAMS=1 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$skip" -c 1 \
      -f -o "gpurun_out/${TAG}_ncu_$name" python "$@" > "gpurun_out/${TAG}_ncu_$name.log" 2>&1
}
cap grad 'k_grad$' 0 tools/profile_c3.py 1
cap assemble 'k_assemble' 0 tools/profile_c3.py 1
cap lauum 'k_lauum' 0 tools/profile_c3.py 1
cap panel10 'k_potrf_panel' 10 tools/profile_c3.py 1
cap diag10 'k_potrf_diag' 10 tools/profile_c3.py 1
cap trtri12 'k_trtri_row' 12 tools/profile_c3.py 1
# 3b. the dataflow kernels on one n = 4000 matrix in flight
cap flow_potrf_n4000 'k_potrf_flow' 1 tools/longstay_one.py 1 4000 2
cap flow_trtri_n4000 'k_trtri_flow' 1 tools/longstay_one.py 1 4000 2 1
# 4. the bench lines themselves (never under a profiler)
python bench.py --impl reference --steps 4 --warmup 1 > "gpurun_out/${TAG}_bench_reference_arm.json" 2> /dev/null
python bench.py > "gpurun_out/${TAG}_bench_n1.json" 2> "gpurun_out/${TAG}_bench_n1.err"
python tools/bench_longstay.py > "gpurun_out/${TAG}_longstay_n4000.json" 2> /dev/null
python tools/bench_latency.py 2> /dev/null | tail -1 > "gpurun_out/${TAG}_latency.json"
python tools/bench_cohort_train.py 256 500 20 300 > "gpurun_out/${TAG}_cohort_train.txt" 2>&1
python tools/bench_predict.py 16 500 > "gpurun_out/${TAG}_online_imputation.json" 2> /dev/null
ls -la gpurun_out | tail -30
# 5. summarise on the box (the raw .ncu-rep files together exceed what gpurun copies back) and
#    hand the committed-evidence files over in gpurun_out/profiles_out/
python tools/refresh_profiles_post_r02.py "${TAG}" > "gpurun_out/${TAG}_post.log" 2>&1
mkdir -p gpurun_out/profiles_out
cp profiles/${TAG}_* profiles/traffic.json gpurun_out/profiles_out/ 2>/dev/null
rm -f gpurun_out/${TAG}_ncu_*.ncu-rep
du -sh gpurun_out
