"""End-to-end cohort training through the shipped front-end (main_cohort_train): synthetic C2-shape
cohort written in the reference's file formats, phase A (random initialisations) + phase B
(lock-step SCG).  usage: python tools/bench_cohort_train.py [patients] [n] [inits] [iters]"""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import expfiles, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
patients = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
inits = int(sys.argv[3]) if len(sys.argv) > 3 else 20
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30
Q, D, R = 5, 24, 8
top = tempfile.mkdtemp()
pats = {f"p{k}": synth.make_patient(D, n, seed=k) for k in range(patients)}
cfg = expfiles.write_experiment(top, Q, D, R, list(range(1, D + 1)), pats, prior_index=0, random_init_num=inits,
                                top_iteration_num=iters)
t0 = time.perf_counter()
out = subprocess.run([os.path.join(ROOT, "medgp_b200", "host", "main_cohort_train"), "--cfg", cfg, "--pans",
                      os.path.join(top, "data", "cohort.txt")], capture_output=True, text=True, check=True)
if os.environ.get("MEDGP_SCG_TRACE"):
    print(out.stderr)
out = out.stdout
print("\n".join(l for l in out.splitlines() if "phase" in l or "Finish" in l or "shard" in l))
print(f"wall {time.perf_counter() - t0:.2f} s")
