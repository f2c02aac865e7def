"""Kernel clustering / mode-kernel estimation (medgp_b200/kernclust.py):
the glue against the outputs of the reference's own medgpc/clustering code
(tests/golden/clustering.json, produced by tests/golden/make_golden_clustering.py), the KDE oracle
against scipy, and -- on the GPU -- the batched KDE kernel against the oracle and the whole
train -> cluster -> test pipeline through the shipped front-ends."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from medgp_b200 import expfiles, kernclust, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "clustering.json")))


def golden_inputs():
    from make_golden_clustering import fitted_thetas
    Q, D, R, count = GOLD["Q"], GOLD["D"], GOLD["R"], GOLD["count"]
    pans = np.array([f"p{i}" for i in range(count)])
    return Q, D, R, pans, fitted_thetas(Q, D, R, count, GOLD["seed"])


def oracle_modes(sets):
    from oracle import oracle_kde
    return [oracle_kde.kde_mode(v) for v in sets]


def test_kde_oracle_against_scipy():
    """the density the oracle restates from statsmodels is the Gaussian KDE scipy computes when
    given the same bandwidth; Silverman's rule on a hand-checked sample"""
    from scipy import stats
    from oracle import oracle_kde
    rng = np.random.default_rng(0)
    for n in (5, 37, 400):
        x = np.concatenate([rng.normal(0.0, 1.0, n), rng.normal(4.0, 0.3, n // 2)])
        h = oracle_kde.bw_silverman(x)
        ref = stats.gaussian_kde(x, bw_method=h / np.std(x, ddof=1)).evaluate(x)
        assert np.abs(oracle_kde.kde_density(x, x) - ref).max() <= 1e-12 * ref.max()
        assert h == kernclust.bw_silverman(x)
    x = np.array([1.0, 2.0, 3.0, 4.0, 100.0])      # IQR/1.349 = 2/1.349 < std: the IQR branch
    assert abs(oracle_kde.bw_silverman(x) - 0.9 * (2.0 / 1.349) * 5 ** -0.2) < 1e-15
    x = np.array([1.0, 1.0, 1.0, 1.0, 2.0])        # IQR = 0: falls back on the standard deviation
    assert abs(oracle_kde.bw_silverman(x) - 0.9 * np.std(x, ddof=1) * 5 ** -0.2) < 1e-15
    with pytest.raises(RuntimeError):
        oracle_kde.bw_silverman(np.ones(4))


def check_against_golden(kde_modes=None, ctx=None, tol=1e-10, tmp="."):
    Q, D, R, pans, hyp = golden_inputs()
    comp_pan, comp_qidx, comp_feature = kernclust.extract_LMC_SM_feature(pans, hyp, Q, D, R)
    assert comp_pan.tolist() == GOLD["comp_pan"] and comp_qidx.tolist() == GOLD["comp_qidx"]
    assert np.abs(comp_feature[0] - np.array(GOLD["feature_rows"]["0"])).max() <= 1e-14
    assert np.abs(comp_feature[-1] - np.array(GOLD["feature_rows"]["last"])).max() <= 1e-14
    assert abs(comp_feature.sum() - GOLD["feature_checksum"][0]) <= 1e-10
    assert abs(np.abs(comp_feature).sum() - GOLD["feature_checksum"][1]) <= 1e-10
    exp_param = {"kernel": "LMC-SM", "Q": Q, "D": D, "R": R, "exp_kernel_dir": os.path.join(str(tmp), "kernel")}
    for g in GOLD["modes"]:
        mode = kernclust.output_mode_LMC_SM(0, exp_param, pans, hyp, comp_pan, comp_qidx, g["cluster_num"],
                                            np.array(g["assign"]), g["name"], ctx=ctx, kde_modes=kde_modes)
        want = np.array(g["mode_hyp"])
        assert mode.shape == want.shape == (D + g["cluster_num"] * (D * R + 2 + D),)
        # A comes out of an SVD: its columns are defined up to sign; compare what the kernel sees,
        # B_q = A_q A_q^T + diag(kappa_q), and everything else entry by entry
        nQ = g["cluster_num"]
        Bm, Bw = kernclust.compute_B_matrix(nQ, D, R, mode), kernclust.compute_B_matrix(nQ, D, R, want)
        for q in range(nQ):
            assert np.abs(Bm[q] - Bw[q]).max() <= tol * max(1.0, np.abs(Bw[q]).max())
        rest = np.r_[0:D, D + nQ * D * R: D + nQ * (D * R + 2 + D)]
        assert np.abs(mode[rest] - want[rest]).max() <= tol * max(1.0, np.abs(want[rest]).max())
        written = kernclust.read_double_from_bin(os.path.join(str(tmp), "kernel", "fold0", f"{g['name']}_mode_param.bin"))
        assert np.array_equal(written, mode)
        assert int(open(os.path.join(str(tmp), "kernel", "fold0", f"{g['name']}_mode_mixture_num.txt")).read()) == nQ


def test_mode_estimation_glue_against_reference_golden(tmp_path):
    check_against_golden(kde_modes=oracle_modes, tmp=tmp_path)


def test_no_cpu_path_for_the_kde():
    Q, D, R, pans, hyp = golden_inputs()
    cp, cq, cf = kernclust.extract_LMC_SM_feature(pans, hyp, Q, D, R)
    with pytest.raises(RuntimeError):
        kernclust.output_mode_LMC_SM(0, {"Q": Q, "D": D, "R": R, "exp_kernel_dir": "/tmp/none"}, pans, hyp, cp, cq, 1,
                                     np.zeros(len(cp), dtype=int), "x")


def test_gmm_clustering_front_end():
    """run_clustering_top on component features: 'gmm' = sklearn GaussianMixture + BIC over 1..Q
    clusters as the reference runs it (cluster.py:23-46), 'None' = one cluster"""
    rng = np.random.default_rng(1)
    feats = []
    for _ in range(30):
        feats.append(kernclust.compute_sm_feature(1.0 / rng.uniform(20, 28), (1.0 / (2 * np.pi * rng.uniform(60, 70))) ** 2))
        feats.append(kernclust.compute_sm_feature(1.0 / rng.uniform(60, 70), (1.0 / (2 * np.pi * rng.uniform(6, 8))) ** 2))
    feats = np.array(feats)
    num, assign = kernclust.run_clustering_top("gmm", feats, max_cluster_num=3, init_num=2, random_state=0)
    assert 1 <= num <= 3 and assign.shape == (len(feats),) and len(np.unique(assign)) == num
    # two well separated blobs in two dimensions: BIC picks two clusters and they are the blobs
    blobs = np.r_[rng.normal(0, 0.1, (80, 2)), rng.normal(3, 0.1, (80, 2))]
    num2, assign2 = kernclust.run_clustering_top("gmm", blobs, max_cluster_num=3, init_num=2, random_state=0)
    assert num2 == 2 and len(set(assign2[:80])) == 1 and len(set(assign2[80:])) == 1 and assign2[0] != assign2[-1]
    num, assign = kernclust.run_clustering_top("None", feats)
    assert num == 1 and (assign == 0).all()


@pytest.mark.gpu
def test_gpu_kde_modes_against_oracle():
    from medgp_b200 import api
    from oracle import oracle_kde
    rng = np.random.default_rng(2)
    sets = [rng.normal(0, 1, 7), rng.lognormal(0, 1, 300), np.r_[rng.normal(-2, 0.2, 2000), rng.normal(3, 1.0, 1500)],
            rng.uniform(0, 1, 257), np.array([0.3, 0.9])]
    ctx = api.Context(2, 3, 2, workspace_bytes=1 << 28)
    bw = [kernclust.bw_silverman(v) for v in sets]
    modes, dens = ctx.kde_mode(sets, bw, want_density=True)
    for v, h, m, d in zip(sets, bw, modes, dens):
        d0 = oracle_kde.kde_density(v, v, h)
        assert np.abs(d - d0).max() <= 1e-12 * d0.max()
        assert abs(m - oracle_kde.kde_mode(v)) <= 1e-12 * max(1.0, abs(m))
    with pytest.raises(api.MedgpError):
        ctx.kde_mode([np.ones(3)], [0.0])
    ctx.close()


@pytest.mark.gpu
def test_gpu_mode_estimation_against_reference_golden(tmp_path):
    from medgp_b200 import api
    ctx = api.Context(GOLD["Q"], GOLD["D"], GOLD["R"], workspace_bytes=1 << 28)
    check_against_golden(ctx=ctx, tmp=tmp_path)
    ctx.close()


@pytest.mark.gpu
def test_train_cluster_test_pipeline_on_one_box(tmp_path):
    """main_cohort_train -> kernel_clustering_top -> main_cohort_test with the files each step writes"""
    from medgp_b200 import api
    Q, D, R = 2, 3, 2
    host = os.path.join(ROOT, "medgp_b200", "host")
    pats = {f"p{k}": synth.make_patient(D, n, seed=700 + k, T=120.0) for k, n in enumerate([60, 75, 50, 66, 58, 71])}
    top = str(tmp_path)
    cfg = expfiles.write_experiment(top, Q, D, R, [1, 3, 4], pats, prior_index=0, random_init_num=6, top_iteration_num=25)
    setup = json.load(open(cfg))
    setup["cohort_id_list"] = "cohort.txt"
    json.dump(setup, open(cfg, "w"), indent=4)
    subprocess.run([os.path.join(host, "main_cohort_train"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt")],
                   check=True, capture_output=True, timeout=900)
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    mode = kernclust.kernel_clustering_top(cfg, fold=-1, algorithm="None", ctx=ctx)
    ctx.close()
    assert mode.shape == (D + Q * (D * R + 2 + D) - (Q - 1) * (D * R + 2 + D),) and np.isfinite(mode).all()
    # main_cohort_test reads kernel/fold<F>/: the all-patients estimate serves fold 0 here
    os.makedirs(os.path.join(top, "kernel", "fold0"), exist_ok=True)
    for f in ("None_mode_param.bin", "None_mode_mixture_num.txt"):
        os.replace(os.path.join(top, "kernel", "all", f), os.path.join(top, "kernel", "fold0", f))
    subprocess.run([os.path.join(host, "main_cohort_test"), "--cfg", cfg, "--pans", os.path.join(top, "data", "cohort.txt"),
                    "--fold", "0", "--kernclust-alg", "None"], check=True, capture_output=True, timeout=900)
    for pan, (m, x, y) in pats.items():
        for mode_name in ("mean_wo_update", "mean_w_update"):
            pred = expfiles.read_double_bin(os.path.join(top, "test", f"test_{mode_name}_pred_{pan}.bin"))
            assert len(pred) == len(x) and np.isfinite(pred).all()
