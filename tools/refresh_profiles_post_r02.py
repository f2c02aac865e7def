"""Turn the raw output of tools/refresh_profiles_r02.sh (gpurun_out/<tag>_*) into the committed
evidence under profiles/: launch list + per-kernel summary of a reduced C3 step, per-stage DRAM
traffic (profiles/traffic.json, read by bench.py for roofline.traffic), one summary per
ncu --set full capture, the SASS opcode table, and the bench lines.
usage: python tools/refresh_profiles_post_r02.py [tag]"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*cmd):
    return subprocess.run([sys.executable, *cmd], capture_output=True, text=True, cwd=ROOT).stdout


for name in ("launches_c3_step.csv", "dram_traffic_c3_step.csv", "profile_c3.json", "bench_n1.json",
             "bench_reference_arm.json", "longstay_n4000.json", "latency.json", "cohort_train.txt",
             "online_imputation.json", "bench_n2.json", "bench_n4.json", "bench_n8.json"):
    f = os.path.join(SRC, f"{TAG}_{name}")
    if os.path.exists(f) and os.path.getsize(f) > 0:
        shutil.copy(f, os.path.join(DST, f"{TAG}_{name}"))
        print("copied", name)

ll = os.path.join(DST, f"{TAG}_launches_c3_step.csv")
if os.path.exists(ll):
    open(os.path.join(DST, f"{TAG}_launches_c3_step.summary.txt"), "w").write(run("tools/launch_summary.py", ll))

STAGE_OF = {"k_potrf_panel": "potrf", "k_potrf_diag": "potrf", "k_potrf_step": "potrf", "k_syrk_update": "potrf",
            "void k_grad<5>": "grad", "k_grad_finish": "grad", "k_lauum": "lauum", "void k_assemble<5>": "assemble",
            "k_trtri_row": "trtri", "k_trtri_update": "trtri"}
tr = os.path.join(DST, f"{TAG}_dram_traffic_c3_step.csv")
pc = os.path.join(DST, f"{TAG}_profile_c3.json")
if os.path.exists(tr) and os.path.exists(pc):
    meta = json.loads(open(pc).read().strip().splitlines()[-1])
    rows = list(csv.DictReader(l for l in open(tr) if not l.startswith("==")))
    per = collections.defaultdict(float)
    launches = collections.defaultdict(set)
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        if k in STAGE_OF:
            per[STAGE_OF[k]] += float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
            launches[STAGE_OF[k]].add(r["ID"])
    traffic = {"source": f"profiles/{TAG}_dram_traffic_c3_step.csv: dram__bytes_read.sum + dram__bytes_write.sum of every launch of one "
                         f"step of the reduced C3 cohort (tools/profile_c3.py: {meta['patients']} patients x {meta['inits']} theta, "
                         "same size distribution), summed per stage and divided by the step's sum of n^2; bench.py scales it to its own "
                         "shard and launch count",
               "sum_n2_of_sample": meta["sum_n2_per_step"]}
    for st, b in per.items():
        traffic[st] = {"dram_bytes_per_n2": b / meta["sum_n2_per_step"], "dram_bytes_sample_step": b,
                       "launches_sample_step": len(launches[st])}
    reps = glob.glob(os.path.join(SRC, f"{TAG}_ncu_*.ncu-rep"))
    for rep in reps:
        name = os.path.basename(rep)[len(TAG) + 5:-len(".ncu-rep")]
        for line in run("tools/ncu_summary.py", rep).splitlines():
            if "sm__pipe_fp64_cycles_active" in line and name in ("grad", "assemble"):
                traffic[f"{name}_fp64_pipe_pct"] = round(float(line.split()[-2]), 1)
            if "sm__pipe_tensor_cycles_active" in line and name not in ("grad", "assemble"):
                traffic[f"{name}_tensor_pipe_pct"] = round(float(line.split()[-2]), 1)
    json.dump(traffic, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
    print("traffic.json", {k: v for k, v in traffic.items() if k != "source"})

for rep in sorted(glob.glob(os.path.join(SRC, f"{TAG}_ncu_*.ncu-rep"))):
    base = os.path.basename(rep)[:-len(".ncu-rep")]
    txt = run("tools/ncu_summary.py", rep) + "\n-- hottest SASS lines (tools/ncu_hot.py) --\n" + run("tools/ncu_hot.py", rep, "2.0")
    txt += "\n-- executed instruction mix (tools/ncu_instmix.py) --\n" + "\n".join(run("tools/ncu_instmix.py", rep).splitlines()[:14]) + "\n"
    open(os.path.join(DST, base + ".summary.txt"), "w").write(txt)
    print("summarised", base)
open(os.path.join(DST, f"{TAG}_sass_opcodes.txt"), "w").write(run("tools/sass_summary.py"))
