cap() {  # name regex skip
  MEDGP_GRAPHS=0 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/r2x_ncu_$1" python tools/longstay_one.py 1 4000 2 > "gpurun_out/r2x_ncu_$1.log" 2>&1
}
cap syrkA 'k_syrk_update' 81
cap panel 'k_potrf_panel' 80
cap diag 'k_potrf_diag' 80
ls -la gpurun_out/r2x*
