"""Writers/readers for MedGP's experiment layout: exp_setup.json,
hyp_bound.txt, per-patient feature files, feature stats, raw-double result files.  Byte
compatible with medgpc/util/config.py:5-105 and medgpc/util/binaryIO.py, so the reference's
Python side and both sets of executables can read what this writes."""
from __future__ import annotations

import json
import os

import numpy as np

from . import synth

OPT_DEFAULT = {  # scripts/opt_prior0.json
    "flag_use_sum_obj": 0, "random_init_num": 1000, "random_seed": 718, "top_iteration_num": 1000,
    "iteration_num_per_update": 30, "online_learn_rate": 0.00001, "online_momentum": 0.9,
    "lower_bound_noise": 0.15, "upper_bound_noise": 0.4, "lower_bound_a": -1.5, "upper_bound_a": 1.5,
    "lower_bound_period": 12, "upper_bound_period": 72, "lower_bound_lengthscale": 6,
    "upper_bound_lengthscale": 72, "lower_bound_lambda": 0.1, "upper_bound_lambda": 0.5,
    "lower_bound_scale": 0.1, "upper_bound_scale": 1.5,
}


def write_hyp_bound(path, D, Q, R, opt):
    """lb/ub pairs, one number per line, theta order (medgpc/util/config.py:38-65)."""
    with open(path, "w") as f:
        def pair(key, count):
            for _ in range(count):
                f.write("{:6.6f}\n".format(opt["lower_bound_" + key]))
                f.write("{:6.6f}\n".format(opt["upper_bound_" + key]))
        pair("noise", D)
        pair("a", Q * D * R)
        pair("period", Q)
        pair("lengthscale", Q)
        pair("lambda", Q * D)


def write_experiment(top, Q, D, R, features, patients, prior_index=0, eta=0.01, beta_lam=0.01, **opt_over):
    """Creates <top>/{config,train,test,kernel/fold0,data}; returns the exp_setup.json path.
    patients: {PAN: (meta, x, y)} with y already z-scored (stats are written as mean 0, std 1)."""
    opt = dict(OPT_DEFAULT)
    opt.update(opt_over)
    for d in ("config", "train", "test", "kernel/fold0", "data"):
        os.makedirs(os.path.join(top, d), exist_ok=True)
    setup = dict(
        data_dir=os.path.join(top, "data"), exp_top_dir=top, exp_train_dir=os.path.join(top, "train"),
        exp_test_dir=os.path.join(top, "test"), exp_kernel_dir=os.path.join(top, "kernel"),
        exp_cfg_dir=os.path.join(top, "config"), hyp_bound_file="hyp_bound.txt", kernel="LMC-SM",
        kernel_index=7, prior="None" if prior_index == 0 else "hier-gamma", prior_index=prior_index,
        Q=Q, D=D, R=R, feature_index="".join(f"{f} " for f in features), eta=eta, beta_lam=beta_lam,
        cv_fold_num=1, cv_assign_file="cv_assign.txt", random_init_num=opt["random_init_num"],
        random_seed=opt["random_seed"], top_iteration_num=opt["top_iteration_num"],
        iteration_num_per_update=opt["iteration_num_per_update"],
        online_learn_rate=opt["online_learn_rate"], online_momentum=opt["online_momentum"])
    cfg = os.path.join(top, "config", "exp_setup.json")
    with open(cfg, "w") as f:
        json.dump(setup, f, indent=4)
    write_hyp_bound(os.path.join(top, "config", "hyp_bound.txt"), D, Q, R, opt)
    for feat in features:
        np.array([0.0, 1.0]).tofile(os.path.join(top, "data", f"feature{feat}_stat.bin"))
    for pan, (meta, x, y) in patients.items():
        os.makedirs(os.path.join(top, "data", pan), exist_ok=True)
        for j, feat in enumerate(features):
            sel = np.asarray(meta) == j
            with open(os.path.join(top, "data", pan, f"feature{feat}.txt"), "w") as fh:
                fh.write(f"{int(sel.sum())}\n")
                for a, b in zip(np.asarray(x)[sel], np.asarray(y)[sel]):
                    fh.write(f"{a:6.6f}\n{b:6.6f}\n")
    with open(os.path.join(top, "data", "cohort.txt"), "w") as f:
        for pan in patients:
            f.write(pan + "\n")
    return cfg


def write_mode_kernel(top, Q, theta, alg="None", fold=0):
    """<exp_kernel_dir>/fold<f>/<alg>_mode_{mixture_num.txt,param.bin} (mode_estimate.py:425-429)."""
    d = os.path.join(top, "kernel", f"fold{fold}")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f"{alg}_mode_mixture_num.txt"), "w") as f:
        f.write(f"{Q}\n")
    np.asarray(theta, dtype=np.float64).tofile(os.path.join(d, f"{alg}_mode_param.bin"))


def read_double_bin(path):
    return np.fromfile(path, dtype=np.float64)


def read_int_txt(path):
    return [int(v) for v in open(path).read().split()]


def reload_patient(top, pan, features):
    """Reads a patient back exactly as c_experiment::get_one_patient_data does (values pass
    through the %6.6f text format, so they differ from the generator's float32 by <= 5e-7)."""
    meta, xs, ys = [], [], []
    for j, feat in enumerate(features):
        vals = open(os.path.join(top, "data", pan, f"feature{feat}.txt")).read().split()
        n = int(float(vals[0]))
        for i in range(n):
            meta.append(j)
            xs.append(np.float32(vals[1 + 2 * i]))
            ys.append(np.float32((float(np.float32(vals[2 + 2 * i])) - 0.0) / 1.0))
    return np.array(meta, dtype=np.int32), np.array(xs, dtype=np.float32), np.array(ys, dtype=np.float32)


__all__ = ["write_experiment", "write_mode_kernel", "write_hyp_bound", "read_double_bin", "read_int_txt",
           "reload_patient", "OPT_DEFAULT", "synth"]
