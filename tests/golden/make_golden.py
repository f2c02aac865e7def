"""Generates tests/golden/golden.json from the UNMODIFIED reference compiled in this container
(oracle/_ref, built by oracle/ref/build_ref.sh from /root/reference).  Run here once; the GPU
box and CI only read the committed JSON.

    python tests/golden/make_golden.py

Contents:
  eval    reference NLML / gradient / prediction (float arithmetic) on seeded synthetic cases
  scg     reference c_optimizer_scg on the analytic objective (tests/cpp/analytic_objective.h)
  varem   reference c_optimizer_varEM on the same objective (EM state, pruning flags)
  init    first random theta vectors of the reference's c_experiment::get_global_hyp
  train   reference main_one_train.o end to end on one tiny patient
  test    reference main_one_test.o end to end (sliding-window imputation, with and without
          online updates) on a patient with shared time stamps: every output file
"""
import importlib.util
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from medgp_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
REFERENCE_ROOT = os.environ.get("MEDGP_REFERENCE", "/root/reference")


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=600,
                          env=oracle.ref_env(), **kw).stdout


def parse_tagged(text):
    out = {}
    for line in text.split("\n"):
        parts = line.split()
        if len(parts) == 2 and parts[0] in ("calls", "loss", "x", "v", "t", "h"):
            out.setdefault(parts[0], []).append(float(parts[1]))
    return out


def write_experiment(tmp, Q, D, R, features, prior_index, n_init, n_iter, patients, **opt_over):
    """exp_setup.json / hyp_bound.txt / data files as medgpc/util/config.py writes them."""
    spec = importlib.util.spec_from_file_location(
        "refconfig", os.path.join(REFERENCE_ROOT, "medgpc", "util", "config.py"))
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    opt = json.load(open(os.path.join(REFERENCE_ROOT, "scripts", "opt_prior0.json")))
    opt["random_init_num"], opt["top_iteration_num"] = n_init, n_iter
    opt.update(opt_over)
    for d in ("config", "train", "test", "kernel/fold0", "data"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    paths = dict(data_dir=os.path.join(tmp, "data"), exp_top_dir=tmp,
                 exp_train_dir=os.path.join(tmp, "train"), exp_test_dir=os.path.join(tmp, "test"),
                 exp_kernel_dir=os.path.join(tmp, "kernel"), exp_cfg_dir=os.path.join(tmp, "config"),
                 hyp_bound_file="hyp_bound.txt")
    cfg.write_medgpc_bound(os.path.join(tmp, "config"), "hyp_bound.txt", D, 7, Q, R, opt)
    cfg.write_medgpc_config_json(os.path.join(tmp, "config", "exp_setup.json"), paths, "LMC-SM", 7,
                                 features, "None" if prior_index == 0 else "hier-gamma", prior_index,
                                 0.01, 0.01, Q, R, opt, 1, "cv_assign.txt")
    for f in features:
        np.array([0.0, 1.0]).tofile(os.path.join(tmp, "data", f"feature{f}_stat.bin"))
    for pan, (meta, x, y) in patients.items():
        os.makedirs(os.path.join(tmp, "data", pan), exist_ok=True)
        for j, f in enumerate(features):
            sel = meta == j
            with open(os.path.join(tmp, "data", pan, f"feature{f}.txt"), "w") as fh:
                fh.write(f"{int(sel.sum())}\n")
                for a, b in zip(x[sel], y[sel]):
                    fh.write(f"{a:6.6f}\n{b:6.6f}\n")
    return os.path.join(tmp, "config", "exp_setup.json")


def main():
    gold = {"note": "outputs of the unmodified reference (g++ -O2, OpenBLAS shim); float arithmetic"}
    # ---- eval
    cases = []
    for (Q, D, R, n, seed) in [(2, 2, 2, 80, 1), (1, 1, 1, 37, 3), (3, 4, 2, 130, 4), (5, 24, 8, 300, 5),
                               (5, 24, 8, 500, 6)]:  # the last one is the C2 shape at full size (T = 8 tiles deep)
        meta, x, y = synth.make_patient(D, n, seed)
        theta = synth.init_hyp_lmc_sm(Q, D, R, 2, seed=718 + seed)[1]
        r = oracle.ref_eval(Q, D, R, meta, x, y, theta, mode=1)
        ms, xs = meta[::7][:5].copy(), (x[::7][:5] + np.float32(0.37)).astype(np.float32)
        pr = oracle.ref_eval(Q, D, R, meta, x, y, theta, mode=2, meta_star=ms, x_star=xs)
        m = len(xs)
        cases.append(dict(Q=Q, D=D, R=R, n=n, seed=seed, theta_seed=718 + seed, nlml=r["nlml"],
                          grad=r["values"].tolist(), star_meta=ms.tolist(), star_x=[float(v) for v in xs],
                          pred_mean=pr["values"][:m].tolist(), pred_var=pr["values"][m:2 * m].tolist()))
    gold["eval"] = cases
    # ---- scg / varem on the analytic objective
    scg = []
    for iters, x0 in [(-5, [0.1, -0.4, 2.5, 1.0, -1.2, 0.7, 3.0, -2.0]), (-30, [0.1, -0.4, 2.5, 1.0, -1.2, 0.7, 3.0, -2.0]),
                      (-25, [5.5, 5.9, -5.7, 5.0]), (-60, [2.0, -3.0, 1.0])]:
        out = parse_tagged(run([os.path.join(REF, "ref_scg"), "scg", str(iters)] + [repr(v) for v in x0]))
        scg.append(dict(max_iteration=iters, x0=x0, calls=int(out["calls"][0]), loss=out["loss"][0], x=out["x"]))
    gold["scg"] = scg
    varem = []
    for iters, sub, Q, D, R, x0 in [(-7, 30, 1, 2, 1, [-1.2, -1.0, 0.8, -0.6, -2.5, -3.0, -1.5, -1.7]),
                                    (-3, 20, 2, 2, 1, [-1.2, -1.0, 0.8, -0.6, 0.3, 1e-9, -2.5, -3.0, -2.2, -2.8, -1.5, -1.7, -1.1, -1.9])]:
        out = parse_tagged(run([os.path.join(REF, "ref_scg"), "varem", str(iters), str(sub), str(Q), str(D), str(R),
                                "0.01", "0.01"] + [repr(v) for v in x0]))
        varem.append(dict(max_iteration=iters, sub_iter=sub, Q=Q, D=D, R=R, x0=x0, calls=int(out["calls"][0]),
                          loss=out["loss"][0], x=out["x"], v=out["v"], t=[int(v) for v in out["t"]]))
    gold["varem"] = varem
    # ---- init + end-to-end train on a tiny patient
    with tempfile.TemporaryDirectory() as tmp:
        Q, D, R, n = 2, 2, 2, 60
        meta, x, y = synth.make_patient(D, n, seed=11)
        cfg = write_experiment(tmp, Q, D, R, [18, 19], 0, 20, 30, {"p0": (meta, x, y)})
        out = parse_tagged(run([os.path.join(REF, "ref_scg"), "init", cfg, "3"]))
        gold["init"] = dict(Q=Q, D=D, R=R, seed=718, count=3, theta=out["h"])
        run([os.path.join(REF, "main_one_train.o"), "--cfg", cfg, "--pan", "p0", "--thread", "1"])
        init_hyp = np.fromfile(os.path.join(tmp, "train", "train_init_hyp_p0.bin"))
        hyp = np.fromfile(os.path.join(tmp, "train", "train_hyp_p0.bin"))
        f_init = oracle.nlml_grad(Q, D, R, meta, x, y, init_hyp, want_grad=False)[0]
        f_fit = oracle.nlml_grad(Q, D, R, meta, x, y, hyp, want_grad=False)[0]
        gold["train"] = dict(Q=Q, D=D, R=R, n=n, patient_seed=11, random_init_num=20, top_iteration_num=30,
                             init_hyp=init_hyp.tolist(), hyp=hyp.tolist(), nlml_init_fp64=f_init, nlml_fit_fp64=f_fit,
                             train_num=int(open(os.path.join(tmp, "train", "train_num_p0.txt")).read()),
                             train_flag=int(open(os.path.join(tmp, "train", "train_flag_p0.txt")).read()))
    # ---- end-to-end test executable: both imputation modes (main_one_test.cpp:447-472 outputs)
    with tempfile.TemporaryDirectory() as tmp:
        Q, D, R, n = 2, 2, 1, 26
        meta, x, y = synth.make_patient(D, n, seed=77, T=100.0)
        x[3] = x[14]   # two features observed at one time stamp
        x[20] = x[7]   # and a second shared stamp
        lr = 1e-3
        cfg = write_experiment(tmp, Q, D, R, [18, 19], 0, 20, 30, {"p0": (meta, x, y)}, online_learn_rate=lr)
        theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=3)[0]
        theta[D + 1] = 0.0   # an A entry at exactly 0 is clamped at test time (c_prior.cpp:118-140)
        with open(os.path.join(tmp, "kernel", "fold0", "None_mode_mixture_num.txt"), "w") as fh:
            fh.write(f"{Q}\n")
        theta.tofile(os.path.join(tmp, "kernel", "fold0", "None_mode_param.bin"))
        run([os.path.join(REF, "main_one_test.o"), "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0",
             "--kernclust-alg", "None"])
        td = os.path.join(tmp, "test")
        modes = {}
        for mode in ("mean_wo_update", "mean_w_update"):
            modes[mode] = dict(
                flag=[int(v) for v in open(os.path.join(td, f"test_{mode}_flag_p0.txt")).read().split()],
                feature=[int(v) for v in open(os.path.join(td, f"test_{mode}_feature_p0.txt")).read().split()],
                ci=[int(v) for v in open(os.path.join(td, f"test_{mode}_ci_p0.txt")).read().split()],
                etime=np.fromfile(os.path.join(td, f"test_{mode}_etime_p0.bin")).tolist(),
                error=np.fromfile(os.path.join(td, f"test_{mode}_error_p0.bin")).tolist(),
                pred=np.fromfile(os.path.join(td, f"test_{mode}_pred_p0.bin")).tolist())
        gold["test"] = dict(Q=Q, D=D, R=R, n=n, patient_seed=77, T=100.0, shared=[[3, 14], [20, 7]], features=[18, 19],
                            online_learn_rate=lr, theta_seed=3, theta=theta.tolist(), modes=modes)
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote tests/golden/golden.json:", {k: (len(v) if isinstance(v, list) else "ok") for k, v in gold.items()})


if __name__ == "__main__":
    main()
