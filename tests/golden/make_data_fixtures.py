"""Regenerates tests/golden/data_*.npz from the reference's ./data directory.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_data_fixtures.py [/root/reference/data]
The arrays are the outputs of autoreparam_b200.data's loaders (which restate the
reference's loaders, models.py:706-760, 860-881, 984-989, 1037-1045); they are
public data sets (UCI German credit, Gelman & Hill radon / election88 / electric
company), stored derived and compressed so tests can run without the raw files.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from autoreparam_b200 import data  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
out = os.path.dirname(os.path.abspath(__file__))


def save(name, d):
    np.savez_compressed(os.path.join(out, "data_%s.npz" % name), **{k: np.asarray(v) for k, v in d.items()})
    print(name, {k: np.asarray(v).shape for k, v in d.items()})


g = data.load_german_credit(src)
# one-hot part is exact in uint8; keep the 8 numeric columns in float32
save("german_credit", {"X": g["X"], "y": g["y"].astype(np.uint8)})
for st in ["PA", "MN"]:
    r = data.load_radon(st, src)
    save("radon_" + st, r)
e = data.load_election(src)
save("election", {"n_state": e["n_state"], "state": e["state"].astype(np.uint8), "female": e["female"].astype(np.uint8),
                  "black": e["black"].astype(np.uint8), "y": e["y"].astype(np.uint8)})
save("electric", data.load_electric(src))
