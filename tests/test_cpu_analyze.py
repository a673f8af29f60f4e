"""Host logic: the results-dir reader (mirror of the reference's analyze.py) on the file names main.py writes."""
import json
import os

import numpy as np

from autoreparam_b200 import analyze


def _vi(elbo, t):
    return {"elbo": elbo, "estimated_elbo_std": 0.25, "variational_fit_time_secs": t, "learning_rate": 0.05,
            "actual_num_variational_steps": 3000, "initial_step_size": [[0.1], [0.2]],
            "learned_variational_params": {"mu_loc": [0.0], "mu_scale": [1.0]}}


def _hmc(d, ess, lf, t):
    d.update({"ess_min": [ess], "sem_min": [ess / 10], "acceptance_rate": [0.8], "mcmc_time_sec": [t],
              "num_leapfrog_steps": lf})
    return d


def _make(tmp_path):
    base = tmp_path / "radon_PA"
    base.mkdir()
    files = {
        "CP_tied.json": _hmc(_vi(-3652.9, 3.0), 22.0, 4, 10.0),
        "NCP_tied.json": _hmc(_vi(-3656.5, 3.0), 3.0, 4, 10.0),
        "cVIP_eig.json": _hmc(dict(_vi(-3652.8, 6.0), learned_reparam={"m_a": [0.9, 0.2]}), 21.0, 8, 12.0),
        "cVIP_eig_tied.json": _hmc(dict(_vi(-3652.7, 6.0), learned_reparam={"m_a": [0.7, 0.4], "mua_a": 0.1}), 23.0, 8, 12.0),
        "dVIP_eig_tied.json": _hmc(dict(_vi(-3652.9, 6.0), learned_reparam={"m_a": [1.0, 0.0]}), 22.5, 4, 9.0),
    }
    i = {"ess_min": [30.0], "sem_min": [3.0], "acceptance_rate_cp": [0.7], "acceptance_rate_ncp": [0.8],
         "mcmc_time_sec": [20.0], "num_leapfrog_steps": [4]}          # as autoreparam_b200.main appends it
    files["i_tied.json"] = i
    for name, d in files.items():
        (base / name).write_text(json.dumps(d))
    return str(tmp_path)


def _run(argv):
    lines = []
    rc = analyze.main(argv, log=lambda x: lines.append(str(x)))
    return rc, lines


def test_candidates_reconcile_reader_and_writer_names(tmp_path):
    d = _make(tmp_path)
    f = lambda m: os.path.basename(analyze.find_result_file(d, "radon_PA", m))
    assert f("CP") == "CP_tied.json" and f("NCP") == "NCP_tied.json" and f("i") == "i_tied.json"
    assert f("cVIP_exp_tied") == "cVIP_eig_tied.json"      # reader says exp, writer's default lpt is eig
    assert f("cVIP_exp") == "cVIP_eig.json"                 # untied file, never the dVIP / tied one
    (tmp_path / "radon_PA" / "CP.json").write_text(json.dumps(_hmc(_vi(-1.0, 1.0), 1.0, 4, 1.0)))
    assert f("CP") == "CP.json"                             # the reader's literal name wins when present


def test_elbos_ess_reparams_lines(tmp_path):
    d = _make(tmp_path)
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--elbos"])
    assert rc == 0 and lines[0] == " ******  radon_PA  ****** "
    assert "-3652.9000 +/- 0.25   : CP" in lines and "-3652.7000 +/- 0.25   : cVIP_exp_tied" in lines
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--ess"])
    assert "[22.0] +/- [2.2] : CP (4 leapfrog steps)" in lines
    assert "[30.0] +/- [3.0] : i (4 leapfrog steps)" in lines           # appended list form of num_leapfrog_steps
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--ess", "--normalize_times"])
    cp = [l for l in lines if l.endswith("x/1.00x CP time per VI/MCMC step)") and ": CP (" in l]
    # 22 ESS per 1000 grads x 4 x 10000 grads = 880 effective samples in 3 + 10 s
    assert cp and cp[0].startswith("880.0 +/- 88.0 in 13.0s (3.0s VI + 10.0s MCMC): CP (4 leapfrog steps, 1.00x/1.00x")
    il = [l for l in lines if ": i (" in l][0]
    assert il.startswith("2400.0 +/- 240.0 in 26.0s (6.0s VI + 20.0s MCMC): i (4 leapfrog steps, 2.00x/1.00x")
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--reparams"])
    assert "   cVIP_exp_tied" in lines and any(l.startswith("       m_a: [0.7 0.4]") for l in lines)


def test_validate(tmp_path):
    d = _make(tmp_path)
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--validate"])
    assert rc == 0 and lines == [" ******  radon_PA  ****** ok"]
    bad = json.loads((tmp_path / "radon_PA" / "NCP_tied.json").read_text())
    bad["ess_min"].append(1.0)          # appended lists out of step
    del bad["elbo"]
    (tmp_path / "radon_PA" / "NCP_tied.json").write_text(json.dumps(bad))
    rc, lines = _run(["--results_dir", d, "--model", "radon_PA", "--validate"])
    assert rc == 1 and any("VI keys missing: elbo" in l for l in lines) and any("differ in length" in l for l in lines)
    rc, lines = _run(["--results_dir", d, "--model", "all", "--validate"])   # other model directories absent: skipped
    assert rc == 1 and sum("******" in l for l in lines) == 1


def test_missing_files_are_reported_not_raised(tmp_path):
    rc, lines = _run(["--results_dir", str(tmp_path), "--model", "8schools_data", "--ess", "--elbos"])
    assert rc == 0 and any("no results file" in l for l in lines)
    assert np.isfinite(0.0)
