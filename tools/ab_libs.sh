#!/bin/bash
# A/B of alternative builds of libmedgp_cuda.so on a reduced C3 cohort: per-stage times (profile
# mode, one stream) and end-to-end wall time.  usage: tools/ab_libs.sh out_prefix lib1 lib2 ...
out=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  MEDGP_LIB=$lib python tools/profile_c3.py 2 ${AB_PATIENTS:-768} 5 > ${out}_${name}_stages.json 2> ${out}_${name}.err
  MEDGP_LIB=$lib python tools/bench_c3.py ${AB_PATIENTS:-768} 5 3 > ${out}_${name}_e2e.json 2>> ${out}_${name}.err
done
python - "$out" "$@" <<'P'
import json, sys, os
out = sys.argv[1]
for lib in sys.argv[2:]:
    name = os.path.basename(lib)[:-3]
    try:
        s = json.load(open(f"{out}_{name}_stages.json")); e = json.load(open(f"{out}_{name}_e2e.json"))
        st = s["stage_ms_per_step"]
        print(f"{name:28s} " + " ".join(f"{k}={v:7.1f}" for k, v in st.items() if v > 0.5) + f" | sum={sum(st.values()):7.1f} e2e={e['s_per_step']*1e3:7.1f} ms")
    except Exception as ex:
        print(name, "FAILED", ex)
P
