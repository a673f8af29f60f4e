import cProfile, pstats, sys, os, io
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench
from autoreparam_b200 import graphs, inference, models
class A: features=25; method="NCP"
raw, D, a, b = bench.workload(A)
mc = models.from_data("german_credit_lognormalcentered", raw)
C, L, S = 16384, 4, 1000
z0, sigma_q = bench.init_states(D, C, 0)
target = graphs.TargetGraph(mc, "NCP", a, b, False)
kw = dict(num_leapfrog_steps=L, num_samples=S, num_burnin_steps=500, num_adaptation_steps=400, device=torch.device("cuda", 0))
inference.hmc(target, mc, mc.split(sigma_q), z0, seed=1, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(3):
    inference.hmc(target, mc, mc.split(sigma_q), z0, seed=2 + i, **kw)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
