"""GPU: the shipped front-ends (medgp_b200/host/*, linked against libmedgp_cuda.so) against the
same sources linked against the oracle (oracle/_build/*), and the CUDA path against the
committed reference outputs (tests/golden/golden.json)."""
import json
import os
import subprocess

import numpy as np
import pytest

from medgp_b200 import expfiles, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "medgp_b200", "host")
BUILD = os.path.join(ROOT, "oracle", "_build")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def run(args, timeout=900):
    return subprocess.run(args, check=True, capture_output=True, text=True, timeout=timeout).stdout


@pytest.mark.parametrize("idx", range(len(GOLD["eval"])))
def test_cuda_path_vs_reference_golden(idx):
    """GPU (FP64) against the unmodified reference (float): the reference's own noise floor"""
    from medgp_b200 import api
    c = GOLD["eval"][idx]
    meta, x, y = synth.make_patient(c["D"], c["n"], c["seed"])
    theta = synth.init_hyp_lmc_sm(c["Q"], c["D"], c["R"], 2, seed=c["theta_seed"])[1]
    ctx = api.Context(c["Q"], c["D"], c["R"], workspace_bytes=1 << 30)
    sid = ctx.add_series(meta, x, y)
    f, g, st = ctx.nlml_grad([sid], theta[None], True)
    assert st[0] == 0
    assert abs(f[0] - c["nlml"]) <= 2e-6 * abs(c["nlml"])
    gref = np.array(c["grad"])
    assert np.abs(g[0] - gref).max() <= 5e-5 * np.abs(gref).max()
    m = len(c["star_x"])
    mean, var, _ = ctx.predict([sid], theta[None], [0, m], c["star_meta"], np.array(c["star_x"], dtype=np.float32))
    assert np.abs(mean - np.array(c["pred_mean"])).max() <= 2e-4
    assert np.abs(var - np.array(c["pred_var"])).max() <= 2e-4
    ctx.close()


@pytest.mark.parametrize("prior_index,budget", [(0, 12), (2, 2)])
def test_train_front_ends_gpu_vs_oracle_build(tmp_path, oracle, prior_index, budget):
    Q, D, R = 2, 3, 2
    pats = {f"p{k}": synth.make_patient(D, n, seed=200 + k) for k, n in enumerate([70, 130, 45])}
    top_g, top_o = str(tmp_path / "gpu"), str(tmp_path / "orc")
    kw = dict(prior_index=prior_index, random_init_num=8, top_iteration_num=budget, iteration_num_per_update=10)
    cfg_g = expfiles.write_experiment(top_g, Q, D, R, [1, 3, 4], pats, **kw)
    cfg_o = expfiles.write_experiment(top_o, Q, D, R, [1, 3, 4], pats, **kw)
    run([os.path.join(HOST, "main_cohort_train"), "--cfg", cfg_g, "--pans", os.path.join(top_g, "data", "cohort.txt")])
    run([os.path.join(BUILD, "main_cohort_train"), "--cfg", cfg_o, "--pans", os.path.join(top_o, "data", "cohort.txt")])
    run([os.path.join(HOST, "main_one_train"), "--cfg", cfg_g, "--pan", "p0", "--thread", "1"])  # overwrites p0 (same result)
    for pan in pats:
        assert expfiles.read_int_txt(os.path.join(top_g, "train", f"train_flag_{pan}.txt")) == [1]
        a = expfiles.read_double_bin(os.path.join(top_g, "train", f"train_init_hyp_{pan}.bin"))
        b = expfiles.read_double_bin(os.path.join(top_o, "train", f"train_init_hyp_{pan}.bin"))
        assert np.array_equal(a, b)
        a = expfiles.read_double_bin(os.path.join(top_g, "train", f"train_hyp_{pan}.bin"))
        b = expfiles.read_double_bin(os.path.join(top_o, "train", f"train_hyp_{pan}.bin"))
        if prior_index == 0:
            # a dozen chained evaluations: GPU and oracle agree to ~1e-12 per evaluation
            assert np.abs(a - b).max() <= 1e-7 * max(1.0, np.abs(b).max())
        else:
            # hundreds of chained evaluations through a chaotic line search:
            # compare the objective reached, not theta
            m, x, y = expfiles.reload_patient(top_g, pan, [1, 3, 4])
            fa = oracle.nlml_grad(Q, D, R, m, x, y, a, want_grad=False)[0]
            fb = oracle.nlml_grad(Q, D, R, m, x, y, b, want_grad=False)[0]
            assert abs(fa - fb) <= 0.02 * abs(fb)


def test_test_front_end_gpu_vs_oracle_build(tmp_path):
    Q, D, R = 2, 2, 1
    meta, x, y = synth.make_patient(D, 40, seed=5, T=120.0)
    x[3] = x[14]          # time stamps shared by two and three observations: the leave-one-out
    x[20] = x[31]         # path inside a time stamp, in both modes
    x[21] = x[31]
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=3)[0]
    outs = {}
    for tag, bindir in (("gpu", HOST), ("orc", BUILD)):
        top = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [18, 19], {"p0": (meta, x, y)}, online_learn_rate=1e-3)
        expfiles.write_mode_kernel(top, Q, theta)
        run([os.path.join(bindir, "main_one_test"), "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0",
             "--kernclust-alg", "None"])
        outs[tag] = {n: expfiles.read_double_bin(os.path.join(top, "test", f"test_{n}_pred_p0.bin"))
                     for n in ("mean_wo_update", "mean_w_update")}
        outs[tag]["ci"] = expfiles.read_int_txt(os.path.join(top, "test", "test_mean_wo_update_ci_p0.txt"))
    for n in ("mean_wo_update", "mean_w_update"):
        assert len(outs["gpu"][n]) == 40
        assert np.abs(outs["gpu"][n] - outs["orc"][n]).max() <= 1e-6
    assert outs["gpu"]["ci"] == outs["orc"]["ci"]


def test_cohort_test_front_end_gpu_vs_oracle_build(tmp_path):
    """main_cohort_test on the GPU (one factorisation per patient, one call per shard) against
    the same source on the oracle backend (one refit per observation)."""
    Q, D, R = 2, 3, 2
    pats = {f"p{k}": synth.make_patient(D, n, seed=400 + k, T=150.0) for k, n in enumerate([60, 131, 47, 90])}
    for m, x, y in pats.values():
        x[5] = x[17]
        x[6] = x[17]
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=9)[0]
    outs = {}
    for tag, bindir in (("gpu", HOST), ("orc", BUILD)):
        top = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats)
        expfiles.write_mode_kernel(top, Q, theta)
        run([os.path.join(bindir, "main_cohort_test"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
             "--fold", "0", "--kernclust-alg", "None"])
        outs[tag] = {pan: (expfiles.read_double_bin(os.path.join(top, "test", f"test_mean_wo_update_pred_{pan}.bin")),
                           expfiles.read_int_txt(os.path.join(top, "test", f"test_mean_wo_update_ci_{pan}.txt")))
                     for pan in pats}
    for pan, (m, x, y) in pats.items():
        assert len(outs["gpu"][pan][0]) == len(x)
        assert np.abs(outs["gpu"][pan][0] - outs["orc"][pan][0]).max() <= 1e-6
        assert outs["gpu"][pan][1] == outs["orc"][pan][1]


def test_one_factorisation_imputation_vs_per_observation_refits_on_the_gpu(tmp_path):
    """Both front-ends on the GPU with and without MEDGP_NO_ONLINE=1: the one-factorisation paths
    (whole patient without updates, one time stamp with updates) against the reference's literal
    one-fit-per-observation procedure run through the same library."""
    Q, D, R = 2, 3, 2
    pats = {f"p{k}": synth.make_patient(D, n, seed=600 + k, T=150.0) for k, n in enumerate([70, 45])}
    for m, x, y in pats.values():
        x[5] = x[17]
        x[6] = x[17]
        x[30] = x[29]
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=12)[0]
    outs = {}
    for tag in ("online", "refit"):
        top = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats, online_learn_rate=1e-3)
        expfiles.write_mode_kernel(top, Q, theta)
        env = dict(os.environ, MEDGP_NO_ONLINE="1" if tag == "refit" else "0")
        subprocess.run([os.path.join(HOST, "main_one_test"), "--cfg", cfg, "--pan", "p0", "--thread", "1", "--fold", "0",
                        "--kernclust-alg", "None"], check=True, capture_output=True, timeout=900, env=env)
        subprocess.run([os.path.join(HOST, "main_cohort_test"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
                        "--fold", "0", "--kernclust-alg", "None"], check=True, capture_output=True, timeout=900, env=env)
        outs[tag] = {(pan, mode): expfiles.read_double_bin(os.path.join(top, "test", f"test_{mode}_pred_{pan}.bin"))
                     for pan in pats for mode in ("mean_wo_update", "mean_w_update")
                     if not (mode == "mean_w_update" and pan != "p0")}
    for key, a in outs["online"].items():
        b = outs["refit"][key]
        assert len(a) == len(pats[key[0]][1])
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max())   # float-rounded outputs of FP64 results
