"""Turn the raw output of tools/refresh_profiles.sh (gpurun_out/<tag>_*) into the committed
evidence under profiles/: launch list + summary, per-launch DRAM traffic (traffic.json, read by
bench.py for roofline.traffic), one summary per ncu --set full capture, and the bench lines.
usage: python tools/refresh_profiles_post.py <tag>"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*cmd):
    return subprocess.run([sys.executable, *cmd], capture_output=True, text=True, cwd=ROOT).stdout


for name in ("launches_c2_step.csv", "dram_traffic_c2_step.csv", "bench_n1.json", "bench_reference_arm.json",
             "online_imputation.json", "cholesky_n4000.json", "c3_ragged.json", "latency.json"):
    f = os.path.join(SRC, f"{TAG}_{name}")
    if os.path.exists(f) and os.path.getsize(f) > 0:
        shutil.copy(f, os.path.join(DST, f"{TAG}_{name}"))
        print("copied", name)

ll = os.path.join(DST, f"{TAG}_launches_c2_step.csv")
if os.path.exists(ll):
    open(os.path.join(DST, f"{TAG}_launches_c2_step.summary.txt"), "w").write(run("tools/launch_summary.py", ll))

# per-launch DRAM traffic by stage
tr = os.path.join(DST, f"{TAG}_dram_traffic_c2_step.csv")
if os.path.exists(tr):
    rows = list(csv.DictReader(l for l in open(tr) if not l.startswith("==")))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows:
        per[r["Kernel Name"].split("(")[0]][r["ID"]] += float(r["Metric Value"].replace(",", ""))
    stage_of = {"k_potrf_panel": "potrf", "k_potrf_diag": "diag", "void k_grad<5>": "grad", "k_lauum": "lauum",
                "void k_assemble<5>": "assemble", "k_trtri_row": "trtri"}
    traffic = {stage_of[k]: sum(v.values()) / len(v) for k, v in per.items() if k in stage_of}
    old = {}
    try:
        old = json.load(open(os.path.join(DST, "traffic.json")))
    except (OSError, ValueError):
        pass
    traffic["grad_fp64_pipe_pct"] = old.get("grad_fp64_pipe_pct")
    traffic["_note"] = (f"avg dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu on tools/profile_step.py "
                        f"(C2 step, 256 x n=500), from profiles/{TAG}_dram_traffic_c2_step.csv")
    reps = glob.glob(os.path.join(SRC, f"{TAG}_ncu_grad.ncu-rep"))
    if reps:
        for line in run("tools/ncu_summary.py", reps[0]).splitlines():
            if "sm__pipe_fp64_cycles_active" in line:
                traffic["grad_fp64_pipe_pct"] = round(float(line.split()[-2]), 1)
    json.dump(traffic, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
    print("traffic.json", traffic)

for rep in sorted(glob.glob(os.path.join(SRC, f"{TAG}_ncu_*.ncu-rep"))):
    base = os.path.basename(rep)[:-len(".ncu-rep")]
    txt = run("tools/ncu_summary.py", rep) + "\n-- hottest SASS lines (tools/ncu_hot.py) --\n" + run("tools/ncu_hot.py", rep, "2.0")
    txt += "\n-- executed instruction mix (tools/ncu_instmix.py) --\n" + "\n".join(run("tools/ncu_instmix.py", rep).splitlines()[:14]) + "\n"
    open(os.path.join(DST, base + ".summary.txt"), "w").write(txt)
    print("summarised", base)
