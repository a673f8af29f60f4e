"""GPU: the CUDA fp64 check build against golden vectors produced by the reference's own
models.py / program_transformations.py (tests/golden/make_reference_golden.py), and the
drop-in CLI end to end (VI -> HMCtuning -> HMC, cVIP -> dVIP) with the reference's file layout."""
import json
import os

import numpy as np
import pytest

from autoreparam_b200 import engine, main as arp_main
from tests import common

pytestmark = pytest.mark.gpu
RULES = ["CP", "NCP", "VIP_a", "VIP_ab", "dVIP"]


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(common.GOLDEN, "reference_logjoint.npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.mark.parametrize("rule", RULES)
@pytest.mark.parametrize("model", common.MODELS)
def test_cuda_fp64_matches_reference_model_code(gold, model, rule):
    key = "%s/%s" % (model, rule)
    if key + "/lp" not in gold:
        pytest.skip("the reference itself cannot evaluate %s" % key)
    mc = common.model_config(model, "PA")
    Z, a, b = gold[key + "/z"], gold[key + "/a"], gold[key + "/b"]
    lp, g, xc = engine.log_joint_grad(mc, Z, a, b, precision="f64")
    np.testing.assert_allclose(lp, gold[key + "/lp"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(xc, gold[key + "/centered"], rtol=1e-10, atol=1e-10)
    lp32, _, xc32 = engine.log_joint_grad(mc, Z, a, b, precision="f32")
    tol = 2e-4 if model == "time_series" else 1e-5
    assert common.rel_err(lp32, gold[key + "/lp"]).max() < tol
    assert common.rel_err(xc32, gold[key + "/centered"]).max() < 1e-5


def _run(args):
    arp_main.main(args)


def test_cli_end_to_end_8schools(tmp_path, capsys):
    rd = str(tmp_path / "8schools_")
    base = ["--model=8schools", "--results_dir=" + rd, "--num_optimization_steps=400", "--num_mc_samples=64",
            "--num_samples=300", "--num_burnin_steps=200", "--num_adaptation_steps=150", "--num_chains=64", "--seed=3"]
    for method in ("CP", "NCP", "cVIP"):
        _run(base + ["--inference=VI", "--method=" + method])
    out = capsys.readouterr().out
    assert "Loading model 8schools with dataset ." in out and "finished optimization with elbo" in out
    assert "step 0 elbo" in out
    names = sorted(os.listdir(rd))
    assert names == ["CP_tied.json", "NCP_tied.json", "cVIP_eig_tied.json"]
    cp = json.load(open(os.path.join(rd, "CP_tied.json")))
    assert set(cp) == {"elbo", "variational_fit_time_secs", "actual_num_variational_steps", "estimated_elbo_std",
                       "learning_rate", "initial_step_size", "learned_reparam", "learned_variational_params"}
    assert cp["actual_num_variational_steps"] == 400 and cp["learned_reparam"] is None
    assert len(cp["initial_step_size"]) == 3 and len(cp["initial_step_size"][2]) == 8
    assert set(cp["learned_variational_params"]) == {"mu_loc", "mu_scale", "log_tau_loc", "log_tau_scale",
                                                     "theta_loc", "theta_scale"}
    cv = json.load(open(os.path.join(rd, "cVIP_eig_tied.json")))
    assert set(cv["learned_reparam"]) == {"mu_a", "log_tau_a", "theta_a"} and len(cv["learned_reparam"]["theta_a"]) == 8
    ncp = json.load(open(os.path.join(rd, "NCP_tied.json")))
    assert ncp["elbo"] > cp["elbo"] - 0.5 and -34 < ncp["elbo"] < -30
    # skip-if-exists (main.py:238-242)
    _run(base + ["--inference=VI", "--method=CP"])
    assert "Already ran experiment VI-CP on model 8schools" in capsys.readouterr().out
    # dVIP reads the cVIP file (main.py:162-179)
    _run(base + ["--inference=VI", "--method=dVIP"])
    dv = json.load(open(os.path.join(rd, "dVIP_eig_tied.json")))
    assert all(set(np.ravel(v)) <= {0.0, 1.0} for v in dv["learned_reparam"].values())
    # tuning runs, then HMC picks the best L (main.py:292-294, 316-329)
    for L in (2, 4):
        _run(base + ["--inference=HMCtuning", "--method=NCP", "--num_leapfrog_steps=%d" % L])
    _run(base + ["--inference=HMCtuning", "--method=NCP", "--num_leapfrog_steps=4"])   # de-duplicated
    ncp = json.load(open(os.path.join(rd, "NCP_tied.json")))
    assert [r["num_leapfrog_steps"] for r in ncp["tuning_runs"]] == [2, 4]
    assert set(ncp["tuning_runs"][0]) == {"num_leapfrog_steps", "ess_min", "sem_min", "acceptance_rate", "mcmc_time",
                                          "num_samples", "num_burnin_steps"}
    _run(base + ["--inference=HMC", "--method=NCP", "--num_chains_to_save=3"])
    out = capsys.readouterr().out
    assert "Number of leaprog steps is set to" in out and "ESS per 1000 gradients:" in out
    ncp = json.load(open(os.path.join(rd, "NCP_tied.json")))
    for k in ("ess_min", "sem_min", "acceptance_rate", "mcmc_time_sec", "num_leapfrog_steps"):
        assert isinstance(ncp[k], list) and len(ncp[k]) == 1
    assert 40 < ncp["acceptance_rate"][0] <= 100 and ncp["ess_min"][0] > 0
    ess = np.load(os.path.join(rd, "NCP_tied_ess.npz"))
    assert ess["theta"].shape == (64, 8) and ess["mu"].shape == (64,)
    tr = np.load(os.path.join(rd, "NCP_tied_traces.npz"))
    assert tr["theta"].shape == (300, 3, 8) and tr["mu"].shape == (300, 3)
    assert os.path.exists(os.path.join(rd, "NCP_tied_ess.txt"))
    # the centred posterior mean of mu for eight schools is about 4.4
    assert 2.0 < tr["mu"].mean() < 7.0
    # HMC with the learned (cVIP) and discretised (dVIP) parameterisations
    _run(base + ["--inference=HMC", "--method=cVIP", "--num_leapfrog_steps=4"])
    _run(base + ["--inference=HMC", "--method=dVIP", "--num_leapfrog_steps=4"])
    assert json.load(open(os.path.join(rd, "cVIP_eig_tied.json")))["ess_min"][0] > 0
    # interleaved CP / NCP (--method=i): needs tuning runs for both CP and NCP (main.py:452-480)
    for L in (2, 4):
        _run(base + ["--inference=HMCtuning", "--method=CP", "--num_leapfrog_steps=%d" % L])
    _run(base + ["--inference=HMC", "--method=i", "--num_chains_to_save=2"])
    out = capsys.readouterr().out
    assert "ESS: " in out
    il = json.load(open(os.path.join(rd, "i_tied.json")))
    assert set(il) == {"initial_step_size_ncp", "initial_step_size_cp", "num_leapfrog_steps", "ess_min", "sem_min",
                       "acceptance_rate_cp", "acceptance_rate_ncp", "mcmc_time_sec"}
    assert il["ess_min"][0] > 0 and 30 < il["acceptance_rate_ncp"][0] <= 100
    assert np.load(os.path.join(rd, "i_tied_traces.npz"))["theta"].shape == (300, 2, 8)
    with pytest.raises(Exception, match="Run VI first"):
        _run(["--model=8schools", "--results_dir=" + str(tmp_path / "empty"), "--inference=HMC", "--method=CP",
              "--num_leapfrog_steps=2"])
    # the results-dir reader (mirror of analyze.py) finds the writer's file names and validates the directory
    from autoreparam_b200 import analyze
    lines = []
    rc = analyze.main(["--results_dir", str(tmp_path), "--model", "8schools_", "--elbos", "--ess", "--validate"],
                      log=lambda x: lines.append(str(x)))
    assert rc == 0 and " ******  8schools_  ****** ok" in lines
    assert any(l.endswith(": NCP") and "+/-" in l for l in lines)                      # ELBO line
    assert any(l.endswith(": NCP (%d leapfrog steps)" % ncp["num_leapfrog_steps"][0]) for l in lines)
    assert any(l.endswith(": i (%d leapfrog steps)" % il["num_leapfrog_steps"][0]) for l in lines)
    assert any("ess_per_sec" in l and "rhat_max" in l for l in lines)


def test_cli_streaming_statistics_radon(tmp_path, capsys):
    """--stream_window: the run keeps no [S, C, D] trace (the mode BASELINE configs[4] needs: 65 536 chains x 10 003
    coordinates) and still writes the reference's result files; the ESS it reports agrees with the stored-trace run
    of the same seed wherever the window resolves the autocorrelation."""
    rd = str(tmp_path / "radon_MN")
    base = ["--model=radon", "--dataset=MN", "--results_dir=" + rd, "--num_optimization_steps=300", "--num_mc_samples=64",
            "--num_samples=400", "--num_burnin_steps=200", "--num_adaptation_steps=150", "--num_chains=48", "--seed=5",
            "--method=NCP"]
    # the radon loader needs the reference's srrs2.dat; the fixture holds the loaded arrays instead
    import autoreparam_b200.models as M
    real_loader = M.load_raw
    M.load_raw = lambda model, dataset=None, data_dir=None: common.raw_data(model, dataset or "MN")
    try:
        _run(base + ["--inference=VI"])
        _run(base + ["--inference=HMC", "--num_leapfrog_steps=4"])
        full = json.load(open(os.path.join(rd, "NCP_tied.json")))
        with np.load(os.path.join(rd, "NCP_tied_ess.npz")) as f:     # read now: the second run rewrites the file
            ess_full = {k: f[k] for k in f.files}
        _run(base + ["--inference=HMC", "--num_leapfrog_steps=4", "--stream_window=64"])
    finally:
        M.load_raw = real_loader
    both = json.load(open(os.path.join(rd, "NCP_tied.json")))
    assert len(both["ess_min"]) == 2 and len(full["ess_min"]) == 1          # HMC appends (main.py:375-391)
    assert both["acceptance_rate"][1] == pytest.approx(both["acceptance_rate"][0], abs=1e-9)   # same seed, same chains
    ess_stream = np.load(os.path.join(rd, "NCP_tied_ess.npz"))
    assert ess_stream["m"].shape == ess_full["m"].shape == (48, 85)
    ratio = ess_stream["m"] / ess_full["m"]
    ok = np.isfinite(ratio)
    assert ok.mean() > 0.99
    r = ratio[ok]
    # the window (64 lags) resolves most series exactly.  Where it is too short the estimate can only come out larger;
    # where fp32 round-off moves the first negative autocorrelation by one lag (rho ~ 0 there) the two estimators
    # truncate at different lags and differ by a few per cent in either direction.
    assert np.median(np.abs(r - 1)) < 5e-3, np.median(np.abs(r - 1))
    assert np.quantile(r, 0.02) > 1 - 5e-3 and r.min() > 0.8, (np.quantile(r, 0.02), r.min())
    assert both["ess_min"][1] >= both["ess_min"][0] * 0.98
