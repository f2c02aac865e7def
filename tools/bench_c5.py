"""C5 (BASELINE.json configs[4]): online test-time imputation (main_one_test's sliding window) over a
cohort of synthetic patients with fixed "fitted" hyper-parameters, both modes, through the shipped
front-end main_cohort_test (one GPU), with the reference's main_one_test.o timed beside it on a
bounded sample.
usage: python tools/bench_c5.py [patients] [n_min] [n_max] [ref_patients] [ref_n]
prints one JSON object."""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from medgp_b200 import expfiles, synth  # noqa: E402

Q, D, R = 5, 24, 8
FEATURES = list(range(1, D + 1))


def make_cohort(top, sizes, seed0):
    pats = {f"p{k}": synth.make_patient(D, int(n), seed=seed0 + k, T=240.0 * n / 500.0) for k, n in enumerate(sizes)}
    cfg = expfiles.write_experiment(top, Q, D, R, FEATURES, pats)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=718)[2]
    expfiles.write_mode_kernel(top, Q, theta)
    return cfg, pats


def run_ours(patients=1024, n_min=200, n_max=500):
    sizes = np.random.default_rng(5).integers(n_min, n_max + 1, patients)
    top = tempfile.mkdtemp(prefix="medgp_c5_")
    try:
        t0 = time.perf_counter()
        cfg, pats = make_cohort(top, sizes, 50000)
        t_write = time.perf_counter() - t0
        t0 = time.perf_counter()
        out = subprocess.run([os.path.join(ROOT, "medgp_b200", "host", "main_cohort_test"), "--cfg", cfg, "--pans",
                              os.path.join(top, "data", "cohort.txt"), "--fold", "0", "--kernclust-alg", "None"],
                             capture_output=True, text=True, check=True).stdout
        wall = time.perf_counter() - t0
        wo = re.search(r"without updates: (\d+) predictions from one factorisation per patient in ([\d.e+-]+) s, (\d+) by per-observation refits; elapsed time = ([\d.e+-]+)", out)
        w = re.search(r"with updates: (\d+) predictions and (\d+) hyper-parameter updates \((\d+) reset\) in (\d+) lock-step super-steps, (\d+) predictions by the single-patient routine; elapsed time = ([\d.e+-]+)", out)
        bd = re.search(r"of which: windows ([\d.e+-]+) s, window uploads ([\d.e+-]+) s, NLML\+gradient calls ([\d.e+-]+) s, SGD steps ([\d.e+-]+) s, imputation calls ([\d.e+-]+) s", out)
        n_obs = int(sizes.sum())
        # every observation is imputed once per mode
        done = all(len(expfiles.read_double_bin(os.path.join(top, "test", f"test_{m}_pred_p{k}.bin"))) == sizes[k]
                   for m in ("mean_wo_update", "mean_w_update") for k in (0, patients // 2, patients - 1))
        return {
            "workload": f"C5: {patients} synthetic patients, n ~ U{{{n_min}..{n_max}}} (total {n_obs} observations), D=24 Q=5 R=8, "
                        "every observation imputed from its past (+ same-stamp observations), one GPU, through main_cohort_test",
            "patients": patients, "observations": n_obs, "complete": bool(done),
            "wo_update": {"predictions": int(wo.group(1)) + int(wo.group(3)), "refits": int(wo.group(3)),
                          "seconds": float(wo.group(4)), "predictions_per_s": n_obs / float(wo.group(4)),
                          "gpu_call_seconds": float(wo.group(2))},
            "w_update": {"predictions": int(w.group(1)) + int(w.group(5)), "updates": int(w.group(2)), "resets": int(w.group(3)),
                         "super_steps": int(w.group(4)), "single_patient_routine": int(w.group(5)), "seconds": float(w.group(6)),
                         "predictions_per_s": n_obs / float(w.group(6)),
                         "seconds_breakdown": dict(zip(("windows", "window_uploads", "nlml_grad_calls", "sgd_steps", "imputation_calls"),
                                                       (float(v) for v in bd.groups()))) if bd else None},
            "front_end_wall_seconds": wall, "input_files_write_seconds": t_write, "unit": "predictions/s",
        }
    finally:
        shutil.rmtree(top, ignore_errors=True)


def run_reference(ref_patients=None, ref_n=150):
    """main_one_test.o (both modes, as it always runs them) on `ref_patients` patients of ref_n points, one
    single-thread process per patient, all concurrently."""
    exe = os.path.join(ROOT, "oracle", "_ref", "main_one_test.o")
    if not os.path.exists(exe):
        return None
    from oracle import oracle
    ref_patients = ref_patients or (os.cpu_count() or 1)
    top = tempfile.mkdtemp(prefix="medgp_c5ref_")
    try:
        cfg, pats = make_cohort(top, [ref_n] * ref_patients, 60000)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe, "--cfg", cfg, "--pan", pan, "--thread", "1", "--fold", "0", "--kernclust-alg", "None"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=oracle.ref_env()) for pan in pats]
        for pr in procs:
            pr.wait()
        wall = time.perf_counter() - t0
        preds = 2 * ref_n * ref_patients  # both modes
        return {"predictions_per_s_both_modes": preds / wall, "seconds": wall, "cores": ref_patients,
                "sample": f"{ref_patients} concurrent single-thread main_one_test.o processes (unmodified reference, g++ -O2 + OpenBLAS), "
                          f"one patient of n={ref_n} each, both modes = {preds} predictions; the reference's cost per prediction grows "
                          "with n (O(n^3) refits, O(P n^2) gradients), so on the C5 cohort (n up to 500) it is slower than this"}
    finally:
        shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    res = run_ours(*(a[:3])) if a else run_ours()
    ref = run_reference(*(a[3:5])) if len(a) > 3 else run_reference()
    res["reference_sample"] = ref
    if ref:
        res["both_modes_predictions_per_s"] = 2 * res["observations"] / (res["wo_update"]["seconds"] + res["w_update"]["seconds"])
    print(json.dumps(res))
