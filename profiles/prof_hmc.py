"""Short HMC run for ncu captures (profiles/README.md has the commands).
usage: python profiles/prof_hmc.py [engine] [chains] [num_results] [features]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from autoreparam_b200 import data, engine, models  # noqa: E402

eng = int(sys.argv[1]) if len(sys.argv) > 1 else 0
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
S = int(sys.argv[3]) if len(sys.argv) > 3 else 40
F = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = data.synthetic_german_credit(f=F)
mc = models.from_data("german_credit_lognormalcentered", raw)
D = mc.num_coords
rng = np.random.default_rng(0)
z0 = torch.as_tensor((0.1269 * rng.standard_normal((C, D))).astype(np.float32), device="cuda")
for rep in range(2):
    out = engine.hmc_run(mc, z0, np.full(D, 0.1269), np.zeros(D), np.zeros(D), num_leapfrog_steps=4, num_results=S,
                         num_burnin_steps=20, num_adaptation_steps=20, seed=rep, engine=eng, want_final=False)
torch.cuda.synchronize()
print("accept", float(out["is_accepted"].float().mean()), "transitions", out["num_transitions"])
