"""CPU: host logic, loaders, file layout, C-ABI exports, multi-process (gloo) plumbing."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from autoreparam_b200 import _lib, data, distributed, graphs, main as arp_main, models, util
from oracle import oracle as O
from tests import common

REF_DATA = "/root/reference/data"


def test_german_credit_fixture_checksums():
    """SURVEY 8c loader pins: 1000 x 62, sum X = 14000, sum X^2 = 20993, rank 49, sum y = 700."""
    d = common.raw_data("german_credit_lognormalcentered")
    X, y = d["X"], d["y"]
    assert X.shape == (1000, 62) and X.dtype == np.float32
    assert abs(X.sum() - 14000) < 1e-2 and abs((X.astype(np.float64) ** 2).sum() - 20993) < 1e-1
    assert np.linalg.matrix_rank(X) == 49 and y.sum() == 700
    np.testing.assert_allclose(X[0, :8], [1, -1.23586, -0.744759, 0.918018, 1.046463, 2.765073, 1.026565, -0.428075],
                               atol=2e-6)


def test_grouped_data_fixture_pins():
    r = common.raw_data("radon", "PA")
    assert len(r["y"]) == 2389 and len(r["u"]) == 68 and np.all(np.diff(r["county"]) >= 0)
    e = common.raw_data("election")
    assert len(e["y"]) == 11566 and (e["state"] == 51).sum() == 15
    assert sorted(set(range(1, 52)) - set(e["state"].tolist())) == [2, 12]
    el = common.raw_data("electric")
    assert len(el["y"]) == 192 and (el["pair"] >= 96).sum() == 2 and (el["grade"] >= 4).sum() == 42
    assert (el["grade_pair"] >= 4).sum() == 21
    assert (data.onehot_index(el["grade"], 4) == -1).sum() == 42


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference data not present (GPU box)")
def test_loaders_reproduce_fixtures_from_raw_files():
    g = data.load_german_credit(REF_DATA)
    np.testing.assert_array_equal(g["X"], common.raw_data("german_credit_lognormalcentered")["X"])
    for st in ("PA", "MN"):
        r, f = data.load_radon(st, REF_DATA), common.raw_data("radon", st)
        for k in ("county", "u", "x", "y"):
            np.testing.assert_array_equal(r[k], f[k])
    e = data.load_election(REF_DATA)
    np.testing.assert_array_equal(e["state"], common.raw_data("election")["state"])
    el = data.load_electric(REF_DATA)
    np.testing.assert_array_equal(el["pair"], common.raw_data("electric")["pair"])


def test_model_zoo_matches_oracle_site_tables():
    for m in common.MODELS:
        mc = models.from_data(m, common.raw_data(m))
        raw = common.raw_data(m)
        assert [(n, tuple(s)) for n, s in O.site_table(m, raw)] == mc.sites
        assert mc.num_coords == O.num_coords(m, raw)
        z = np.arange(2 * mc.num_coords, dtype=np.float64).reshape(2, -1)
        np.testing.assert_array_equal(mc.join(mc.split(z)), z)
    with pytest.raises(Exception, match="unknown model"):
        models.get_model_by_name("no_such_model")
    with pytest.raises(NotImplementedError):
        models.get_model_by_name("gp_poisson")


def test_method_to_rule_table():
    mc = models.from_data("8schools", common.raw_data("8schools"))
    assert (graphs.make_cp_graph(mc).a == 1).all() and (graphs.make_cp_graph(mc).b == 1).all()
    assert (graphs.make_ncp_graph(mc).a == 0).all() and (graphs.make_ncp_graph(mc).b == 0).all()
    g = graphs.make_cvip_graph(mc, "eig", tied_pparams=True)
    assert g.learnable and (g.a == 0.5).all() and (g.b == 1).all()          # as written: b = 1
    g2 = graphs.make_cvip_graph(mc, "eig", tied_pparams=True, tied_b_as_written=False)
    assert (g2.b == 0.5).all()
    reparam = {"mu_a": 0.7, "log_tau_a": 0.2, "theta_a": list(np.linspace(0, 1, 8)), "theta_prior_mean": [0] * 8}
    d = graphs.make_dvip_graph(mc, graphs.discretise(reparam))
    assert d.a.tolist() == [1.0, 0.0] + [0, 0, 0, 0, 1, 1, 1, 1] and (d.b == 1).all()
    a, b = graphs.reparam_to_ab(mc, {"mu_a": 0.3, "mu_b": 0.4, "log_tau_a": 1.0, "theta_a": 0.5})
    assert a[0] == 0.3 and b[0] == 0.4 and b[1] == 1.0 and (a[2:] == 0.5).all()


def test_results_file_names_match_reference_scheme():
    """main.py:208-219 defaults: CP_tied.json, NCP_tied.json, cVIP_eig_tied.json, dVIP_eig_tied.json."""
    P = arp_main.build_parser()
    names = {m: arp_main.results_filename(P.parse_args(["--method", m])) for m in ("CP", "NCP", "cVIP", "dVIP", "i")}
    assert names == {"CP": "CP_tied.json", "NCP": "NCP_tied.json", "cVIP": "cVIP_eig_tied.json",
                     "dVIP": "dVIP_eig_tied.json", "i": "i_tied.json"}
    f = P.parse_args(["--method", "cVIP", "--tied_pparams", "False", "--discrete_prior", "--learnable_parameterisation_type", "exp"])
    assert arp_main.results_filename(f) == "cVIP_exp_discrete_prior.json"
    assert arp_main.cvip_path(P.parse_args([]), "r") == os.path.join("r", "cVIP_eig_tied.json")
    d = P.parse_args([])
    assert (d.num_samples, d.num_chains, d.num_burnin_steps, d.num_adaptation_steps, d.num_mc_samples,
            d.num_optimization_steps) == (50000, 100, 10000, 6000, 256, 3000)


def test_save_hmc_results_appends(tmp_path):
    p = str(tmp_path / "CP_tied.json")
    json.dump({"elbo": -1.0}, open(p, "w"))
    arp_main.save_hmc_results(p, ess_min=1.0, sem_min=0.1)
    arp_main.save_hmc_results(p, ess_min=2.0, sem_min=0.2)
    arp_main.save_hmc_results(p, tuning_runs={"num_leapfrog_steps": 4, "ess_min": 3.0})
    r = json.load(open(p))
    assert r["ess_min"] == [1.0, 2.0] and r["elbo"] == -1.0 and r["tuning_runs"][0]["num_leapfrog_steps"] == 4
    assert arp_main.get_best_num_leapfrog_steps_from_tuning_runs(
        [{"num_leapfrog_steps": 2, "ess_min": 1.0}, {"num_leapfrog_steps": 8, "ess_min": 5.0}]) == 8
    arp_main.save_ess(str(tmp_path / "CP_tied"), None, [np.ones((3,)), np.ones((3, 8))], ["mu", "theta"])
    assert set(np.load(str(tmp_path / "CP_tied_ess.npz")).files) == {"mu", "theta"}


def test_util_mirrors():
    rng = np.random.default_rng(0)
    ess = [rng.uniform(1, 5, (6,)), rng.uniform(1, 5, (6, 8))]
    ess[1][2, 3] = np.nan
    assert np.allclose(util.get_min_ess(ess, 6), O.get_min_ess(ess, 6))
    params = {"mu_loc": np.float32(1.0), "mu_scale": np.float32(0.5), "theta_loc": np.zeros(8), "theta_scale": np.ones(8)}
    s = util.variational_inits_from_params(params, ["mu", "theta"], 7, rng=np.random.default_rng(1))
    assert s["mu"].shape == (7,) and s["theta"].shape == (7, 8) and s["theta"].dtype == np.float32
    steps = util.get_approximate_step_size(params, 2)
    assert steps[0] == 0.125 and (steps[1] == 0.25).all()
    x = rng.standard_normal((500, 4, 3))
    r = util.rhat_from_moments(x.mean(0), x.var(0), 500)
    assert np.abs(r - 1).max() < 0.05


def test_shard_ranges_partition_the_chain_axis():
    for C, W in ((100, 8), (16384, 8), (7, 3), (5, 8)):
        parts = [distributed.shard_range(C, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == C
        assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
        assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_c_abi_library_exports_every_declared_symbol():
    """The shared libraries load on a CPU-only box and export what include/*.h declares."""
    root = os.path.dirname(common.GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
    header = open(os.path.join(root, "include", "autoreparam_b200.h")).read()
    declared = set(re.findall(r"\b(arp_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for prec in ("f32", "f64"):
        path = _lib.lib_path(prec)
        if not os.path.exists(path):
            import __graft_entry__
            __graft_entry__.build()
        lib = ctypes.CDLL(path)
        for sym in declared:
            assert hasattr(lib, sym), (prec, sym)
        lib.arp_precision.restype = ctypes.c_char_p
        assert lib.arp_precision().decode() == prec
        lib.arp_kernel_launch_count.restype = ctypes.c_int64
        assert lib.arp_kernel_launch_count() == 0


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "autoreparam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


GLOO_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from autoreparam_b200 import distributed, util
distributed.init_if_needed("gloo")
rank, world = distributed.rank_world()
C, D, S = 11, 5, 400
rng = np.random.default_rng(0)
samples = rng.standard_normal((S, C, D)) + np.arange(C)[None, :, None] * 0.01
ess_all = rng.uniform(1, 9, (C, D))
lo, hi = distributed.shard_range(C, rank, world)
g = distributed.gather_chains(ess_all[lo:hi])
assert np.array_equal(g, ess_all), "gather"
tot = distributed.sum_scalar(float(hi - lo))
assert tot == C
mine = samples[:, lo:hi]
r = distributed.rhat_allreduce(mine.mean(0), mine.var(0), S)
ref = util.rhat_from_moments(samples.mean(0), samples.var(0), S)
assert np.allclose(r, ref, rtol=1e-10), (r, ref)
# the tensor path inference.hmc / main.run_hmc use (device tensors + one all-reduce; gloo tensors here)
import torch
st = distributed.rhat_stats(torch.as_tensor(mine.mean(0)), torch.as_tensor(mine.var(0)))
distributed.allreduce_sum_(st)
assert int(round(float(st[-1]))) == C, "chain count is reduced over the ranks"
r2 = distributed.rhat_from_stats(st, S).numpy()
assert np.allclose(r2, ref, rtol=1e-10), (r2, ref)
acc = torch.tensor([float(hi - lo), 2.0 * (hi - lo)], dtype=torch.float64)
distributed.allreduce_sum_(acc)
assert acc.tolist() == [float(C), 2.0 * C]
distributed.barrier(); distributed.shutdown()
open(os.path.join(sys.argv[2], "ok_%d" % rank), "w").write("ok")
"""


def test_sharded_reductions_world_size_2_gloo(tmp_path):
    """N > 1 host path on CPU: 2 processes, gloo; ragged shards (11 chains over 2 ranks)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import socket
    with socket.socket() as sk:      # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), root, str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


def test_cvip_parameter_maps_tied_and_untied():
    """make_cvip_graph: which learnable parameter every coordinate's (a, b) reads
    (program_transformations.py:486-533, 563-566): tied as written (b = 1), the paper's tie b = a, and the untied mode
    where `a` has the shape of the site's loc and `b` of its scale."""
    from autoreparam_b200 import graphs
    from tests import common
    mc = common.model_config("german_credit_lognormalcentered")
    F, D = 62, 125
    t = graphs.make_cvip_graph(mc, tied_pparams=True, tied_b_as_written=True)
    assert t.num_params == D and (t.a_index == np.arange(D)).all() and (t.b_index == -1).all()
    assert (t.a == 0.5).all() and (t.b == 1.0).all()
    t = graphs.make_cvip_graph(mc, tied_pparams=True, tied_b_as_written=False)
    assert t.num_params == D and (t.b_index == t.a_index).all() and (t.b == 0.5).all()
    lr = graphs.learned_reparam_from_params(t, np.linspace(0.1, 0.9, D))
    assert set(lr) == {"overall_log_scale_a", "overall_log_scale_b", "beta_log_scales_a", "beta_log_scales_b",
                       "beta_a", "beta_b"} and np.array_equal(lr["beta_a"], lr["beta_b"])
    a, b = graphs.reparam_to_ab(mc, lr)
    assert np.allclose(a, np.linspace(0.1, 0.9, D), atol=1e-6) and np.allclose(b, a)
    # untied: beta_log_scales ~ N(loc = overall_log_scale [scalar], scale = ones(F)): ONE a, F b's
    t = graphs.make_cvip_graph(mc, tied_pparams=False)
    keys = [(k, sh) for k, _, sh in t.params]
    assert keys == [("overall_log_scale_a", ()), ("overall_log_scale_b", ()), ("beta_log_scales_a", ()),
                    ("beta_log_scales_b", (F,)), ("beta_a", (F,)), ("beta_b", (F,))]
    assert t.num_params == 2 + 1 + F + 2 * F
    assert len(set(t.a_index[1:1 + F])) == 1 and len(set(t.b_index[1:1 + F])) == F
    lr = graphs.learned_reparam_from_params(t, np.random.default_rng(0).uniform(size=t.num_params))
    assert np.shape(lr["beta_log_scales_a"]) == () and np.shape(lr["beta_log_scales_b"]) == (F,)
    a, b = graphs.reparam_to_ab(mc, lr)
    assert np.allclose(a[1:1 + F], float(lr["beta_log_scales_a"])) and np.allclose(b[1:1 + F], lr["beta_log_scales_b"])
    # gammascale: beta_log_scales is not a Normal site -> never reparameterised, a = b = 1 there
    mg = common.model_config("german_credit_gammascale")
    t = graphs.make_cvip_graph(mg, tied_pparams=True)
    assert (t.a_index[1:1 + F] == -1).all() and (t.a[1:1 + F] == 1.0).all() and t.num_params == 1 + F
    # electric: a ~ N(loc [96, 1], scale = 1. scalar); mua / sigma_y / b have scalar locs and vector scales
    me = common.model_config("electric")
    t = graphs.make_cvip_graph(me, tied_pparams=False)
    shapes = dict((k, sh) for k, _, sh in t.params)
    assert shapes["a_a"] == (96, 1) and shapes["a_b"] == () and shapes["mua_a"] == () and shapes["mua_b"] == (4,)


def test_leapfrog_grid_flag_parsing():
    from autoreparam_b200 import main
    P = main.build_parser()
    assert main._parse_leapfrog_steps(P.parse_args(["--num_leapfrog_steps=4"])) == [4]
    assert main._parse_leapfrog_steps(P.parse_args(["--num_leapfrog_steps", "2,4,8,16"])) == [2, 4, 8, 16]
    assert main._parse_leapfrog_steps(P.parse_args([])) == []
    with pytest.raises(NotImplementedError):
        main.check_supported(P.parse_args(["--model=gp_poisson"]))
    with pytest.raises(NotImplementedError):
        main.check_supported(P.parse_args(["--reparameterise_variational"]))
    main.check_supported(P.parse_args(["--discrete_prior", "--tied_pparams=False"]))


def test_discrete_prior_density_known_values():
    """Oracle restatement of the mixture-of-Laplace prior (main.py:244-253) against a direct evaluation."""
    import torch
    from oracle import oracle as O
    p = torch.tensor([0.0, 0.03, 0.5, 0.97, 1.0], dtype=torch.float64)
    got = O.discrete_prior_logp(p).numpy()
    w = np.exp([0.0, 5.0, 0.0]); w /= w.sum()
    lap = lambda x, m: np.exp(-np.abs(x - m) / 0.1) / 0.2
    ref = np.log(w[0] * lap(p.numpy(), 0.0) + w[1] * 1.0 + w[2] * lap(p.numpy(), 1.0))
    assert np.allclose(got, ref, atol=1e-12)
    assert got[0] > got[2] and abs(got[0] - got[4]) < 1e-12     # mass piles up at 0 and 1, symmetric


GLOO_CLI_WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch
from autoreparam_b200 import distributed, inference, main, models
from tests import common
distributed.init_if_needed("gloo")
rank, world = distributed.rank_world()
seen = {}

def fake_hmc(target, mc, step_size_init, z0, reparam=None, *, num_leapfrog_steps, num_samples, num_burnin_steps,
             num_adaptation_steps, num_chains_to_save=0, seed=0, chain_offset=0, device="cpu", precision="f32",
             return_is_accepted=True, stream_window=0, **kw):
    # stands in for the CUDA engine: deterministic per GLOBAL chain id, so the gathered result can be checked
    C, D = z0.shape
    gid = chain_offset + np.arange(C)
    seen.update(C=C, chain_offset=chain_offset, n_save=num_chains_to_save)
    ess_flat = (100.0 + gid[:, None] + 0.01 * np.arange(D)[None, :]).astype(np.float32)
    mean = torch.as_tensor(np.tile(gid[:, None] * 0.001, (1, D)))
    var = torch.ones((C, D), dtype=torch.float64)
    stats = distributed.rhat_stats(mean, var)
    acc = torch.tensor([float(num_samples * C) * 0.5, 1.0, float(C)], dtype=torch.float64)
    distributed.allreduce_sum_(stats)        # the collectives of the real inference.hmc
    distributed.allreduce_sum_(acc)
    rhat = distributed.rhat_from_stats(stats, num_samples).numpy()
    samples = None
    if num_chains_to_save > 0:
        samples = mc.split(np.zeros((num_samples, num_chains_to_save, D), np.float32))
    return inference.HmcResult(ess=mc.split(ess_flat), is_accepted=None, samples=samples, rhat=rhat,
                               step_mult=np.ones(C), accept_count=np.ones(C), num_transitions=1, ess_flat=ess_flat,
                               accept_stats=tuple(float(v) for v in acc))

inference.hmc = fake_hmc
models.load_raw = lambda model, dataset=None, data_dir=None: common.raw_data(model, dataset or "MN")
rd = os.path.join(sys.argv[2], "8schools_")
os.makedirs(rd, exist_ok=True)
if rank == 0:     # what a VI run leaves behind
    json.dump({"initial_step_size": [0.1, 0.1, [0.1] * 8],
               "learned_variational_params": {"mu_loc": 0.0, "mu_scale": 1.0, "log_tau_loc": 0.0, "log_tau_scale": 1.0,
                                              "theta_loc": [0.0] * 8, "theta_scale": [1.0] * 8}},
              open(os.path.join(rd, "NCP_tied.json"), "w"))
distributed.barrier()
C = 11
main.main(["--model=8schools", "--results_dir=" + rd, "--inference=HMC", "--method=NCP", "--num_leapfrog_steps=4",
           "--num_samples=50", "--num_burnin_steps=10", "--num_adaptation_steps=5", "--num_chains=%d" % C,
           "--num_chains_to_save=3"])
lo, hi = distributed.shard_range(C, rank, world)
assert (seen["C"], seen["chain_offset"]) == (hi - lo, lo), seen          # contiguous shard, global chain ids
assert seen["n_save"] == (3 if rank == 0 else 0)                          # traces live on rank 0
distributed.barrier()
if rank == 0:
    out = json.load(open(os.path.join(rd, "NCP_tied.json")))
    assert abs(out["acceptance_rate"][0] - 50.0) < 1e-9, out["acceptance_rate"]     # over the chains of ALL ranks
    ess = np.load(os.path.join(rd, "NCP_tied_ess.npz"))
    assert ess["theta"].shape == (C, 8)                                             # gathered over the ranks
    want = 1000.0 * (100.0 + np.arange(C)[:, None] + 0.01 * (2 + np.arange(8))[None, :]) / (50 * 4)
    assert np.allclose(ess["theta"], want, rtol=1e-6)
    from autoreparam_b200 import util
    ref = util.rhat_from_moments(np.tile(np.arange(C)[:, None] * 0.001, (1, 10)), np.ones((C, 10)), 50)
    assert abs(out["rhat_max"][0] - float(np.max(ref))) < 1e-9 and len(out["ess_min"]) == 1      # all 11 chains
else:
    assert not os.path.exists(os.path.join(rd, "NCP_tied_ess_rank1.npz"))
# more ranks than chains: refused on every rank before any work (no rank is left waiting in a collective)
import types
try:
    main._check_sharding(types.SimpleNamespace(num_chains=1), world)
    raise SystemExit("expected ValueError")
except ValueError as e:
    assert "smaller than the number of ranks" in str(e)
distributed.barrier(); distributed.shutdown()
open(os.path.join(sys.argv[2], "cli_ok_%d" % rank), "w").write("ok")
"""


def test_cli_hmc_path_world_size_2_gloo(tmp_path):
    """The drop-in driver under torchrun (2 processes, gloo, engine stubbed on the CPU): chains sharded contiguously
    with their global ids, ESS gathered over the ranks, acceptance / R-hat reduced over all ranks, files written by
    rank 0 only, and the more-ranks-than-chains case refused on every rank."""
    script = tmp_path / "cli_worker.py"
    script.write_text(GLOO_CLI_WORKER)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), root, str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert (tmp_path / "cli_ok_0").exists() and (tmp_path / "cli_ok_1").exists()


def test_bench_workload_table():
    """bench.py's per-model bookkeeping: every in-scope model has a roofline bound and an algorithmic flop count that
    follows SURVEY.md 8d; the method -> (a, b) map; the reference arm needs no GPU."""
    import argparse
    import bench as B
    from tests import common
    assert set(B.MODEL_TABLE) >= set(common.MODELS) | {"german_synth", "radon_synth"}
    raw = common.raw_data("radon", "PA")
    D = 3 + len(raw["u"])
    assert B.flop_per_grad(argparse.Namespace(model="radon"), raw, D) == 6.0 * 2389 + 8.0 * 68       # 6N + 8J
    rawg = common.raw_data("german_synth")
    assert B.flop_per_grad(argparse.Namespace(model="german_synth"), rawg, 51) == 1.06e5              # 4NF + 12F + 6N
    rawr = common.raw_data("german_credit_lognormalcentered")
    assert B.flop_per_grad(argparse.Namespace(model="german_credit_lognormalcentered"), rawr, 125) == 2.55e5
    a, b = B.method_ab("NCP", 7)
    assert not a.any() and not b.any()
    a, b = B.method_ab("CP", 7)
    assert a.all() and b.all()
    a, b = B.method_ab("dVIP", 7)
    assert set(a) == {0.0, 1.0} and b.all()
    assert B.MODEL_TABLE["radon_synth"]["bound"] == "hbm" and B.MODEL_TABLE["german_synth"]["bound"] == "tensor"
