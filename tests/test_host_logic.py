"""Host-side logic (optimiser steppers, priors, experiment files, front-ends) exercised on CPU.
The executables under oracle/_build/ are the real host sources linked against the ORACLE behind
the C ABI (oracle/oracle_backend.c) -- a test-only build; the shipped binaries need the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

from medgp_b200 import expfiles, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "oracle", "_build")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


@pytest.fixture(scope="module", autouse=True)
def build_host_on_oracle():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libmedgp_oracle.so", "host_on_oracle"],
                   check=True, stdout=subprocess.DEVNULL)


def run(args, timeout=600):
    return subprocess.run(args, check=True, capture_output=True, text=True, timeout=timeout).stdout


def tagged(text):
    out = {}
    for line in text.split("\n"):
        p = line.split()
        if len(p) == 2 and p[0] in ("calls", "loss", "x", "v", "t", "h", "lp", "dlp"):
            out.setdefault(p[0], []).append(float(p[1]))
    return out


@pytest.mark.parametrize("idx", range(len(GOLD["scg"])))
def test_scg_stepper_follows_reference(idx):
    """same number of objective calls and the same iterate as the reference's c_optimizer_scg
    (differences: summation order of the dot products, ~1e-14 per step)"""
    c = GOLD["scg"][idx]
    out = tagged(run([os.path.join(BUILD, "host_check"), "scg", str(c["max_iteration"])] + [repr(v) for v in c["x0"]]))
    assert int(out["calls"][0]) == c["calls"]
    assert abs(out["loss"][0] - c["loss"]) <= 1e-10 * abs(c["loss"])
    assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-7


@pytest.mark.parametrize("idx", range(len(GOLD["varem"])))
def test_varem_stepper_follows_reference(idx):
    c = GOLD["varem"][idx]
    out = tagged(run([os.path.join(BUILD, "host_check"), "varem", str(c["max_iteration"]), str(c["sub_iter"]),
                      str(c["Q"]), str(c["D"]), str(c["R"]), "0.01", "0.01"] + [repr(v) for v in c["x0"]]))
    # each EM round restarts SCG at an almost converged point, where line-search branches hinge
    # on the last bits of f: allow the evaluation count to differ by a couple of calls
    assert abs(int(out["calls"][0]) - c["calls"]) <= 2
    assert abs(out["loss"][0] - c["loss"]) <= 1e-9 * abs(c["loss"])
    assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-5
    assert np.abs(np.array(out["v"]) - np.array(c["v"])).max() <= 1e-6 * max(1.0, np.abs(c["v"]).max())
    assert [int(v) for v in out["t"]] == c["t"]      # same pruning / prior-type flags


@pytest.mark.parametrize("idx", range(len(GOLD["scg"]) + len(GOLD["varem"])))
def test_device_optimiser_driver_follows_reference(idx):
    """medgp_optimize_on_device (variational-EM rounds, result handling; the session itself is
    the oracle backend's here) on the analytic objective against the reference's optimisers.
    The device state machine proper is tests/test_gpu_optimizer.py."""
    if idx < len(GOLD["scg"]):
        c = GOLD["scg"][idx]
        out = tagged(run([os.path.join(BUILD, "host_check"), "dscg", str(c["max_iteration"])] + [repr(v) for v in c["x0"]]))
        assert int(out["calls"][0]) == c["calls"]
        assert abs(out["loss"][0] - c["loss"]) <= 1e-10 * abs(c["loss"])
        assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-7
    else:
        c = GOLD["varem"][idx - len(GOLD["scg"])]
        out = tagged(run([os.path.join(BUILD, "host_check"), "dvarem", str(c["max_iteration"]), str(c["sub_iter"]),
                          str(c["Q"]), str(c["D"]), str(c["R"]), "0.01", "0.01"] + [repr(v) for v in c["x0"]]))
        assert abs(int(out["calls"][0]) - c["calls"]) <= 2
        assert abs(out["loss"][0] - c["loss"]) <= 1e-9 * abs(c["loss"])
        assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-5
        assert [int(v) for v in out["t"]] == c["t"]


def test_random_initialisation_is_bit_identical(tmp_path):
    g = GOLD["init"]
    Q, D, R = g["Q"], g["D"], g["R"]
    meta, x, y = synth.make_patient(D, 30, seed=1)
    cfg = expfiles.write_experiment(str(tmp_path), Q, D, R, [18, 19], {"p0": (meta, x, y)}, random_init_num=5)
    out = tagged(run([os.path.join(BUILD, "host_check"), "init", cfg, str(g["count"])]))
    assert out["h"] == g["theta"]                                   # C++ host == reference
    py = synth.init_hyp_lmc_sm(Q, D, R, g["count"], seed=g["seed"]).ravel()
    assert np.array_equal(py, np.array(g["theta"]))                 # Python generator == reference


def test_prior_functions(oracle):
    for t, xv, p0, p1 in [(1, 0.3, 0.0, 2.0), (1, -1.2, 0.5, 0.01), (2, -0.3, 0.0, 0.5), (2, 0.0, 0.0, 0.5), (2, 4.0, 1.0, 0.01)]:
        out = tagged(run([os.path.join(BUILD, "host_check"), "prior", str(t), repr(xv), repr(p0), repr(p1)]))
        lp, dlp = oracle.prior(t, xv, p0, p1)
        assert out["lp"][0] == lp and out["dlp"][0] == dlp


def _patients(D, sizes, seed0):
    return {f"p{k}": synth.make_patient(D, n, seed=seed0 + k) for k, n in enumerate(sizes)}


def test_main_one_train_end_to_end_vs_reference_golden(tmp_path, oracle):
    g = GOLD["train"]
    Q, D, R = g["Q"], g["D"], g["R"]
    meta, x, y = synth.make_patient(D, g["n"], seed=g["patient_seed"])
    cfg = expfiles.write_experiment(str(tmp_path), Q, D, R, [18, 19], {"p0": (meta, x, y)},
                                    random_init_num=g["random_init_num"], top_iteration_num=g["top_iteration_num"])
    run([os.path.join(BUILD, "main_one_train"), "--cfg", cfg, "--pan", "p0", "--thread", "1"])
    tr = os.path.join(str(tmp_path), "train")
    assert expfiles.read_int_txt(os.path.join(tr, "train_num_p0.txt")) == [g["train_num"]]
    assert expfiles.read_int_txt(os.path.join(tr, "train_flag_p0.txt")) == [g["train_flag"]]
    init_hyp = expfiles.read_double_bin(os.path.join(tr, "train_init_hyp_p0.bin"))
    assert np.array_equal(init_hyp, np.array(g["init_hyp"]))        # same best random init as the reference
    hyp = expfiles.read_double_bin(os.path.join(tr, "train_hyp_p0.bin"))
    m2, x2, y2 = expfiles.reload_patient(str(tmp_path), "p0", [18, 19])
    f_fit = oracle.nlml_grad(Q, D, R, m2, x2, y2, hyp, want_grad=False)[0]
    # float reference vs FP64 backend diverge along the chaotic SCG path:
    # compare the achieved objective, not theta
    assert f_fit < g["nlml_init_fp64"]
    assert f_fit <= g["nlml_fit_fp64"] + 0.05 * abs(g["nlml_init_fp64"] - g["nlml_fit_fp64"])


@pytest.mark.parametrize("prior_index", [0, 2])
def test_cohort_driver_equals_one_patient_runs(tmp_path, prior_index):
    Q, D, R = 2, 2, 2
    pats = _patients(D, [40, 55, 33], 50)
    pats["tiny"] = (np.array([0, 0, 1], dtype=np.int32), np.array([1, 2, 3], dtype=np.float32),
                    np.array([0.1, 0.2, 0.3], dtype=np.float32))     # a feature with one point: skipped
    top_a, top_b = str(tmp_path / "a"), str(tmp_path / "b")
    kw = dict(prior_index=prior_index, random_init_num=6, top_iteration_num=3 if prior_index == 2 else 25,
              iteration_num_per_update=10)
    cfg_a = expfiles.write_experiment(top_a, Q, D, R, [18, 19], pats, **kw)
    cfg_b = expfiles.write_experiment(top_b, Q, D, R, [18, 19], pats, **kw)
    for pan in pats:
        run([os.path.join(BUILD, "main_one_train"), "--cfg", cfg_a, "--pan", pan, "--thread", "1"])
    run([os.path.join(BUILD, "main_cohort_train"), "--cfg", cfg_b, "--pans", os.path.join(top_b, "data", "cohort.txt")])
    for pan in pats:
        for name in (f"train_num_{pan}.txt", f"train_flag_{pan}.txt"):
            assert open(os.path.join(top_a, "train", name)).read() == open(os.path.join(top_b, "train", name)).read()
        if pan == "tiny":
            assert expfiles.read_int_txt(os.path.join(top_b, "train", f"train_flag_{pan}.txt")) == [0]
            continue
        files = [f"train_init_hyp_{pan}.bin", f"train_hyp_{pan}.bin"] + ([f"train_var_hyp_{pan}.bin"] if prior_index == 2 else [])
        for name in files:
            a = expfiles.read_double_bin(os.path.join(top_a, "train", name))
            b = expfiles.read_double_bin(os.path.join(top_b, "train", name))
            assert np.array_equal(a, b), name       # lock-step batching changes nothing


def test_cohort_loader_equals_patient_loader(tmp_path):
    """sizes-only pass + parallel cohort load (c_experiment::get_cohort_sizes / get_cohort_data)
    give what get_one_patient_data gives patient by patient, including empty features"""
    Q, D, R = 2, 3, 2
    pats = _patients(D, [40, 55, 33, 71], 90)
    pats["gap"] = (np.array([0, 0, 2, 2, 2], dtype=np.int32), np.array([1, 2, 3, 4, 5], dtype=np.float32),
                   np.array([0.1, 0.2, 0.3, -0.4, 0.5], dtype=np.float32))   # feature 1 has no observation
    cfg = expfiles.write_experiment(str(tmp_path), Q, D, R, [1, 3, 4], pats)
    out = run([os.path.join(BUILD, "host_check"), "sizes", cfg] + list(pats))
    ns = [int(line.split()[1]) for line in out.split("\n") if line.startswith("n ")]
    same = [int(line.split()[1]) for line in out.split("\n") if line.startswith("same ")]
    assert ns == [len(p[1]) for p in pats.values()] and same == [1] * len(pats)


def test_cohort_driver_resumes(tmp_path):
    """--resume leaves the patients with train_flag 1 alone and trains the others"""
    Q, D, R = 2, 2, 2
    pats = _patients(D, [40, 55, 33], 70)
    top = str(tmp_path)
    cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], pats, random_init_num=4, top_iteration_num=8)
    cohort = os.path.join(top, "data", "cohort.txt")
    run([os.path.join(BUILD, "main_cohort_train"), "--cfg", cfg, "--pans", cohort])
    first = {pan: expfiles.read_double_bin(os.path.join(top, "train", f"train_hyp_{pan}.bin")) for pan in pats}
    os.remove(os.path.join(top, "train", "train_flag_p1.txt"))
    np.zeros(3).tofile(os.path.join(top, "train", "train_hyp_p0.bin"))       # would be overwritten by a re-train
    out = run([os.path.join(BUILD, "main_cohort_train"), "--cfg", cfg, "--pans", cohort, "--resume"])
    assert "resume: 2 patients" in out
    assert np.array_equal(expfiles.read_double_bin(os.path.join(top, "train", "train_hyp_p0.bin")), np.zeros(3))
    assert np.array_equal(expfiles.read_double_bin(os.path.join(top, "train", "train_hyp_p1.bin")), first["p1"])
    assert expfiles.read_int_txt(os.path.join(top, "train", "train_flag_p1.txt")) == [1]


def test_cohort_sharding_covers_every_patient(tmp_path):
    Q, D, R = 1, 2, 1
    pats = _patients(D, [30, 45, 38, 52, 41], 80)
    top = str(tmp_path)
    cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], pats, random_init_num=3, top_iteration_num=5)
    for s in range(2):
        run([os.path.join(BUILD, "main_cohort_train"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
             "--shard", f"{s}/2"])
    for pan in pats:
        assert expfiles.read_int_txt(os.path.join(top, "train", f"train_flag_{pan}.txt")) == [1]


def test_main_one_test_matches_python_replay(tmp_path, oracle):
    """sliding-window imputation, both modes, against a Python replay of
    main_one_test.cpp:269-444 built on the oracle"""
    Q, D, R = 2, 2, 1
    rng = np.random.default_rng(4)
    meta, x, y = synth.make_patient(D, 26, seed=77, T=100.0)
    x[3] = x[14]                                   # two features observed at the same time stamp
    top = str(tmp_path)
    cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], {"p0": (meta, x, y)}, online_learn_rate=1e-3)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=3)[0]
    theta[D + 1] = 0.0                             # an A entry at exactly 0 is clamped at test time
    expfiles.write_mode_kernel(top, Q, theta)
    run([os.path.join(BUILD, "main_one_test"), "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0",
         "--kernclust-alg", "None"])
    meta, x, y = expfiles.reload_patient(top, "p0", [18, 19])
    for update, name in ((False, "mean_wo_update"), (True, "mean_w_update")):
        best, delta = theta.copy(), np.zeros_like(theta)
        preds, errs, cis, feats = [], [], [], []
        stamps = np.unique(x)
        last = stamps[0]
        for tt, s in enumerate(stamps):
            past = (x < s) & ((np.abs(x - s) <= 72.0) if update else True)
            curr = np.nonzero(x == s)[0]
            if update and tt > 3 and (s - last) > np.float32(5.0 / 60.0):
                last = s
                f, g, st = oracle.nlml_grad(Q, D, R, meta[past], x[past], y[past], best)
                if st >= 0 and past.sum() > 2:
                    for h in range(len(theta)):
                        clamped = D <= h < D + Q * D * R and theta[h] == 0.0
                        if not clamped:
                            delta[h] = 0.9 * delta[h] + 1e-3 * g[h]
                            best[h] -= delta[h]
                else:
                    best, delta = theta.copy(), np.zeros_like(theta)
            for jj in curr:
                others = [k for k in curr if k != jj]
                tm = np.concatenate([meta[past], meta[others]])
                tx = np.concatenate([x[past], x[others]])
                ty = np.concatenate([y[past], y[others]])
                if len(tx) > 0:
                    mu, var, _ = oracle.predict(Q, D, R, tm, tx, ty, best, meta[jj:jj + 1], x[jj:jj + 1])
                    mu32, var32 = np.float32(mu[0]), np.float32(var[0])
                    preds.append(float(mu32))
                    err = float(np.float32(mu32 - y[jj]))
                    cis.append(int(abs(err) <= 1.96 * np.sqrt(float(var32))))
                else:
                    preds.append(0.0)
                    err = float(0.0 - y[jj])
                    cis.append(int(abs(err) <= 1.96 * np.exp(theta[meta[jj]])))
                errs.append(err)
                feats.append([18, 19][meta[jj]])
        td = os.path.join(top, "test")
        assert expfiles.read_int_txt(os.path.join(td, f"test_{name}_flag_p0.txt")) == [1]
        assert expfiles.read_int_txt(os.path.join(td, f"test_{name}_feature_p0.txt")) == feats
        got = expfiles.read_double_bin(os.path.join(td, f"test_{name}_pred_p0.bin"))
        assert len(got) == len(x)                      # every observation is imputed once
        tol = 1e-6 if not update else 1e-5
        assert np.abs(got - np.array(preds)).max() <= tol
        assert np.abs(expfiles.read_double_bin(os.path.join(td, f"test_{name}_error_p0.bin")) - np.array(errs)).max() <= tol
        assert expfiles.read_int_txt(os.path.join(td, f"test_{name}_ci_p0.txt")) == cis


def run_test_executable_against_golden(exe, top, env=None):
    """Runs a main_one_test executable on the committed case of GOLD["test"] and compares every
    output file with what the UNMODIFIED reference main_one_test.o wrote (main_one_test.cpp:447-472),
    both modes.  Tolerance: the reference's float arithmetic (2e-4 on predictions and errors);
    a confidence-interval flag may only differ where the error sits on the interval's edge."""
    g = GOLD["test"]
    Q, D, R = g["Q"], g["D"], g["R"]
    meta, x, y = synth.make_patient(D, g["n"], seed=g["patient_seed"], T=g["T"])
    for a, b in g["shared"]:
        x[a] = x[b]
    cfg = expfiles.write_experiment(top, Q, D, R, g["features"], {"p0": (meta, x, y)},
                                    online_learn_rate=g["online_learn_rate"])
    expfiles.write_mode_kernel(top, Q, np.array(g["theta"]))
    if os.path.basename(exe) == "main_cohort_test":
        args = [exe, "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"), "--fold", "0", "--kernclust-alg", "None"]
    else:
        args = [exe, "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0", "--kernclust-alg", "None"]
    subprocess.run(args, check=True, capture_output=True, text=True, timeout=600, env=env)
    td = os.path.join(top, "test")
    for mode, ref in g["modes"].items():
        assert expfiles.read_int_txt(os.path.join(td, f"test_{mode}_flag_p0.txt")) == ref["flag"]
        assert expfiles.read_int_txt(os.path.join(td, f"test_{mode}_feature_p0.txt")) == ref["feature"]
        pred = expfiles.read_double_bin(os.path.join(td, f"test_{mode}_pred_p0.bin"))
        err = expfiles.read_double_bin(os.path.join(td, f"test_{mode}_error_p0.bin"))
        etime = expfiles.read_double_bin(os.path.join(td, f"test_{mode}_etime_p0.bin"))
        assert len(pred) == g["n"] == len(ref["pred"])
        assert np.abs(pred - np.array(ref["pred"])).max() <= 2e-4
        assert np.abs(err - np.array(ref["error"])).max() <= 2e-4
        assert np.array_equal(etime, np.array(ref["etime"]))
        ci = expfiles.read_int_txt(os.path.join(td, f"test_{mode}_ci_p0.txt"))
        assert sum(int(a != b) for a, b in zip(ci, ref["ci"])) <= 1 and len(ci) == len(ref["ci"])


def test_main_one_test_vs_reference_golden(tmp_path):
    """the host front-ends (on the oracle backend) against the reference's own test executable"""
    run_test_executable_against_golden(os.path.join(BUILD, "main_one_test"), str(tmp_path))
    # the cohort front-end (batched over patients; here a cohort of one) writes the same files
    run_test_executable_against_golden(os.path.join(BUILD, "main_cohort_test"), str(tmp_path / "cohort"))
    # ... and with the reference's literal one-fit-per-observation procedure
    run_test_executable_against_golden(os.path.join(BUILD, "main_one_test"), str(tmp_path / "refit"),
                                       env=dict(os.environ, MEDGP_NO_ONLINE="1"))


REF = os.path.join(ROOT, "oracle", "_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main_one_train_orb.o")),
                               reason="oracle/_ref not built (needs /root/reference once)")


def run_reference_binding_checks(suffix, tmp_path, oracle):
    """The drop-in boundary proven by compilation: the UNMODIFIED reference sources with
    c_inference_cuda (oracle/ref/c_inference_cuda.{h,cpp}) in place of c_inference_prior, linked
    against the C ABI (`suffix` = "cuda": libmedgp_cuda.so, "orb": the oracle backend), against
    the reference's own executables on the same experiment directory."""
    Q, D, R = 2, 3, 2
    pats = {"p0": synth.make_patient(D, 90, seed=321)}
    # "host": this repo's own front-end on the SAME backend as the binding
    host_dir = os.path.join(ROOT, "medgp_b200", "host") if suffix == "cuda" else BUILD
    exes = {"ref": os.path.join(REF, "main_one_train.o"), "bind": os.path.join(REF, f"main_one_train_{suffix}.o"),
            "host": os.path.join(host_dir, "main_one_train")}
    env = oracle.ref_env()
    for prior_index, budget in ((0, 12), (2, 1)):
        kw = dict(prior_index=prior_index, random_init_num=6, top_iteration_num=budget, iteration_num_per_update=10)
        res = {}
        for tag, exe in exes.items():
            top = str(tmp_path / f"{tag}{prior_index}")
            cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats, **kw)
            subprocess.run([exe, "--cfg", cfg, "--pan", "p0", "--thread", "1"], check=True, capture_output=True,
                           text=True, timeout=900, env=env)
            tr = os.path.join(top, "train")
            res[tag] = dict(init=expfiles.read_double_bin(os.path.join(tr, "train_init_hyp_p0.bin")),
                            hyp=expfiles.read_double_bin(os.path.join(tr, "train_hyp_p0.bin")),
                            flag=expfiles.read_int_txt(os.path.join(tr, "train_flag_p0.txt")),
                            num=expfiles.read_int_txt(os.path.join(tr, "train_num_p0.txt")))
            if prior_index == 2:
                res[tag]["var"] = expfiles.read_double_bin(os.path.join(tr, "train_var_hyp_p0.bin"))
        assert res["bind"]["flag"] == res["ref"]["flag"] == [1] and res["bind"]["num"] == res["ref"]["num"]
        assert np.array_equal(res["bind"]["init"], res["ref"]["init"])   # same best random initialisation
        if prior_index == 0:
            # a 12-evaluation budget: the float reference and the FP64 backend have not diverged yet
            assert np.abs(res["bind"]["hyp"] - res["ref"]["hyp"]).max() <= 1e-3
        else:
            # one variational-EM round = 100 chained evaluations WITH the prior terms: the binding
            # inside the reference's optimiser walks the path of this repo's host layer on the
            # same backend (both FP64) ...
            assert np.abs(res["bind"]["hyp"] - res["host"]["hyp"]).max() <= 1e-5
            assert np.abs(res["bind"]["var"] - res["host"]["var"]).max() <= 1e-5
            # ... and reaches the float reference's objective (theta itself is chaotic)
            m, x, y = expfiles.reload_patient(str(tmp_path / "ref2"), "p0", [1, 3, 4])
            fb = oracle.nlml_grad(Q, D, R, m, x, y, res["bind"]["hyp"], want_grad=False)[0]
            fr = oracle.nlml_grad(Q, D, R, m, x, y, res["ref"]["hyp"], want_grad=False)[0]
            assert abs(fb - fr) <= 0.03 * abs(fr)
    # single evaluations through the reference's c_objective_one::compute_objective -> binding
    for c in GOLD["eval"][:4]:
        meta, x, y = synth.make_patient(c["D"], c["n"], c["seed"])
        theta = synth.init_hyp_lmc_sm(c["Q"], c["D"], c["R"], 2, seed=c["theta_seed"])[1]
        path = str(tmp_path / "case.txt")
        oracle.write_case(path, c["Q"], c["D"], c["R"], meta, x, y, theta)
        out = subprocess.run([os.path.join(REF, f"ref_eval_{suffix}"), path, "1", "1", "1"], check=True,
                             capture_output=True, text=True, timeout=600, env=env).stdout
        r = oracle.parse_ref_output(out)
        assert r["ok"] and abs(r["nlml"] - c["nlml"]) <= 2e-6 * abs(c["nlml"])
        gref = np.array(c["grad"])
        assert np.abs(r["values"] - gref).max() <= 5e-5 * np.abs(gref).max()
        # the reference's own GP_Regression::predict on the exported chol_alpha / chol_factor_inv
        oracle.write_case(path, c["Q"], c["D"], c["R"], meta, x, y, theta, np.array(c["star_meta"]),
                          np.array(c["star_x"], dtype=np.float32))
        out = subprocess.run([os.path.join(REF, f"ref_eval_{suffix}"), path, "2", "1", "1"], check=True,
                             capture_output=True, text=True, timeout=600, env=env).stdout
        r = oracle.parse_ref_output(out)
        m = len(c["star_x"])
        assert r["ok"]
        assert np.abs(r["values"][:m] - np.array(c["pred_mean"])).max() <= 2e-4
        assert np.abs(r["values"][m:2 * m] - np.array(c["pred_var"])).max() <= 2e-4
    # the reference's test executable with the binding (factors exported for every held-out fit),
    # against the committed outputs of the reference's own main_one_test.o
    run_test_executable_against_golden(os.path.join(REF, f"main_one_test_{suffix}.o"), str(tmp_path / "test_exe"), env=env)


@needs_ref
def test_reference_binding_on_oracle_backend(tmp_path, oracle):
    run_reference_binding_checks("orb", tmp_path, oracle)


def test_cohort_test_front_end_writes_what_main_one_test_writes(tmp_path):
    """main_cohort_test (one batched online-imputation call per shard) against main_one_test run
    patient by patient: identical test_mean_wo_update_* files, shards cover every patient."""
    Q, D, R = 2, 2, 1
    pats = _patients(D, [22, 35, 28], 60)
    for k, (m, x, y) in enumerate(pats.values()):
        x[2 + k] = x[9 + k]                          # shared time stamps
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=3)[0]
    tops = {}
    for tag in ("cohort", "single"):
        top = tops[tag] = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], pats, online_learn_rate=1e-3)
        expfiles.write_mode_kernel(top, Q, theta)
        if tag == "cohort":
            for s in range(2):
                run([os.path.join(BUILD, "main_cohort_test"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
                     "--fold", "0", "--kernclust-alg", "None", "--shard", f"{s}/2"])
        else:
            for pan in pats:
                run([os.path.join(BUILD, "main_one_test"), "--cfg", cfg, "--pan", pan, "--thread", "1", "--fold", "0",
                     "--kernclust-alg", "None"])
    for pan in pats:
        # without updates: identical files; with updates (lock-step over the shard's patients, two
        # batched calls per time-stamp index) the same numbers up to the float rounding of outputs
        for mode, exact in (("mean_wo_update", True), ("mean_w_update", False)):
            for kind in ("pred", "error", "etime"):
                a = expfiles.read_double_bin(os.path.join(tops["cohort"], "test", f"test_{mode}_{kind}_{pan}.bin"))
                b = expfiles.read_double_bin(os.path.join(tops["single"], "test", f"test_{mode}_{kind}_{pan}.bin"))
                assert len(a) == len(pats[pan][1])
                assert np.array_equal(a, b) if exact else np.abs(a - b).max() <= 1e-6
            for kind in ("ci", "feature", "flag"):
                a = expfiles.read_int_txt(os.path.join(tops["cohort"], "test", f"test_{mode}_{kind}_{pan}.txt"))
                b = expfiles.read_int_txt(os.path.join(tops["single"], "test", f"test_{mode}_{kind}_{pan}.txt"))
                assert a == b


def test_per_observation_fallback_writes_the_same_files(tmp_path):
    """MEDGP_NO_ONLINE=1 forces one fit per observation (the path taken when the one-factorisation
    imputation does not apply); both front-ends must write the same files either way."""
    Q, D, R = 2, 2, 1
    pats = _patients(D, [24, 31], 61)
    for k, (m, x, y) in enumerate(pats.values()):
        x[1 + k] = x[8 + k]
        x[12] = x[8 + k]
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=4)[0]
    tops = {}
    for tag in ("online", "refit"):
        top = tops[tag] = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], pats, online_learn_rate=1e-3)
        expfiles.write_mode_kernel(top, Q, theta)
        env = dict(os.environ, MEDGP_NO_ONLINE="1" if tag == "refit" else "0")
        subprocess.run([os.path.join(BUILD, "main_one_test"), "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0",
                        "--kernclust-alg", "None"], check=True, capture_output=True, env=env)
        subprocess.run([os.path.join(BUILD, "main_cohort_test"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
                        "--fold", "0", "--kernclust-alg", "None"], check=True, capture_output=True, env=env)
    for pan in pats:
        for mode in ("mean_wo_update", "mean_w_update"):
            if mode == "mean_w_update" and pan != "p0":
                continue
            a = expfiles.read_double_bin(os.path.join(tops["online"], "test", f"test_{mode}_pred_{pan}.bin"))
            b = expfiles.read_double_bin(os.path.join(tops["refit"], "test", f"test_{mode}_pred_{pan}.bin"))
            assert len(a) == len(pats[pan][1]) and np.array_equal(a, b)
