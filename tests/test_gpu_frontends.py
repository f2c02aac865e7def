"""GPU: the shipped front-ends (medgp_b200/host/*, linked against libmedgp_cuda.so) against the
same sources linked against the oracle (oracle/_build/*), and the CUDA path against the
committed reference outputs (tests/golden/golden.json)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from medgp_b200 import expfiles, synth

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

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


def test_test_executable_gpu_vs_reference_golden(tmp_path):
    """The shipped main_one_test (CUDA backend) against the committed outputs of the UNMODIFIED
    reference main_one_test.o (tests/golden/make_golden.py), both imputation modes, through the
    one-factorisation paths and through the reference's literal one-fit-per-observation procedure."""
    from test_host_logic import run_test_executable_against_golden
    run_test_executable_against_golden(os.path.join(HOST, "main_one_test"), str(tmp_path / "online"))
    run_test_executable_against_golden(os.path.join(HOST, "main_one_test"), str(tmp_path / "refit"),
                                       env=dict(os.environ, MEDGP_NO_ONLINE="1"))
    # the cohort front-end: both modes batched over patients (here a cohort of one)
    run_test_executable_against_golden(os.path.join(HOST, "main_cohort_test"), str(tmp_path / "cohort"))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "main_one_train_cuda.o")),
                    reason="oracle/_ref not built (needs /root/reference once)")
def test_reference_executable_with_cuda_binding(tmp_path, oracle):
    """oracle/_ref/main_one_train_cuda.o = unmodified reference sources + c_inference_cuda +
    -lmedgp_cuda (oracle/ref/build_ref.sh): train_hyp / train_flag against the reference's own
    main_one_train.o after a 12-evaluation budget, a variational-EM round with prior terms, and
    single evaluations through the reference's compute_objective against the golden vectors."""
    from test_host_logic import run_reference_binding_checks
    run_reference_binding_checks("cuda", tmp_path, oracle)


@pytest.mark.parametrize("attempts", [2])
def test_jitter_success_through_main_one_train(tmp_path, attempts):
    """The jitter-success branch through the training executable: with the first two
    factorisation attempts of every evaluation declared failed (MEDGP_FORCE_FAIL, honoured by the
    CUDA library and by the oracle backend alike) main_one_train must walk the same optimiser
    path on the GPU as on the oracle -- i.e. NLML and gradients of K + 3 sigma^2 agree evaluation
    after evaluation (c_inference_exact.cpp:99-108)."""
    Q, D, R = 2, 3, 2
    pats = {"p0": synth.make_patient(D, 90, seed=321)}
    kw = dict(prior_index=0, random_init_num=6, top_iteration_num=12)
    env = dict(os.environ, MEDGP_FORCE_FAIL=str(attempts))
    hyp = {}
    for tag, exe in (("gpu", os.path.join(HOST, "main_one_train")), ("orc", os.path.join(BUILD, "main_one_train"))):
        top = str(tmp_path / tag)
        cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats, **kw)
        out = subprocess.run([exe, "--cfg", cfg, "--pan", "p0", "--thread", "1"], check=True, capture_output=True,
                             text=True, timeout=900, env=env).stdout
        assert f"jittered {attempts} time(s)" in out or tag == "orc"
        assert expfiles.read_int_txt(os.path.join(top, "train", "train_flag_p0.txt")) == [1]
        hyp[tag] = (expfiles.read_double_bin(os.path.join(top, "train", "train_init_hyp_p0.bin")),
                    expfiles.read_double_bin(os.path.join(top, "train", "train_hyp_p0.bin")))
    assert np.array_equal(hyp["gpu"][0], hyp["orc"][0])
    assert np.abs(hyp["gpu"][1] - hyp["orc"][1]).max() <= 1e-7 * max(1.0, np.abs(hyp["orc"][1]).max())
    # and the forced jitter changed the problem: without it the fitted theta is different
    top = str(tmp_path / "plain")
    cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats, **kw)
    run([os.path.join(BUILD, "main_one_train"), "--cfg", cfg, "--pan", "p0", "--thread", "1"])
    plain = expfiles.read_double_bin(os.path.join(top, "train", "train_hyp_p0.bin"))
    assert np.abs(plain - hyp["orc"][1]).max() > 1e-4


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
