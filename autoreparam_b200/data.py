"""Data loaders for the in-scope models (host side, numpy/pandas only).

Each loader restates the corresponding loader of the reference so that the
arrays handed to the CUDA library are the ones the reference's TF graph would
have seen:

* ``load_german_credit``  -- reference ``models.py:860-881`` + the design-matrix
  assembly inside the model body ``models.py:889-892`` (intercept | standardised
  numerics | one-hot categoricals).
* ``load_radon``          -- reference ``models.py:706-760`` (index-fragile
  uranium mapping restated literally).
* ``load_election``       -- reference ``models.py:984-989`` reading
  ``data/election88.py``.
* ``load_electric``       -- reference ``models.py:1037-1045`` reading
  ``data/electric.py``.
* ``eight_schools`` / ``time_series`` -- literal data of ``models.py:134-137``
  and ``models.py:1098-1113``.

The reference reads ``./data/`` relative to the working directory
(``models.py:46``).  We do the same, but the directory can be overridden with
``ARP_DATA_DIR`` (or the ``data_dir`` argument) so tests can point at fixtures.
"""
from __future__ import annotations

import os
import runpy

import numpy as np

DEFAULT_DATA_DIR = "./data/"


def _data_dir(data_dir=None):
    if data_dir is not None:
        return data_dir
    return os.environ.get("ARP_DATA_DIR", DEFAULT_DATA_DIR)


# --------------------------------------------------------------------------- #
# literal data
# --------------------------------------------------------------------------- #
def eight_schools():
    """Reference ``models.py:134-137``."""
    y = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype=np.float32)
    sigma = np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype=np.float32)
    return {"y": y, "sigma": sigma}


def time_series():
    """Reference ``models.py:1098-1113`` (Mauna-Loa style yearly CO2 series)."""
    x = np.arange(1959, 2019, dtype=np.float32)
    y = np.array([
        315.97, 316.91, 317.64, 318.45, 318.99, 319.62, 320.04, 321.38, 322.16,
        323.04, 324.62, 325.68, 326.32, 327.45, 329.68, 330.18, 331.11, 332.04,
        333.83, 335.4, 336.84, 338.75, 340.11, 341.45, 343.05, 344.65, 346.12,
        347.42, 349.19, 351.57, 353.12, 354.39, 355.61, 356.45, 357.1, 358.83,
        360.82, 362.61, 363.73, 366.7, 368.38, 369.55, 371.14, 373.28, 375.8,
        377.52, 379.8, 381.9, 383.79, 385.6, 387.43, 389.9, 391.65, 393.85,
        396.52, 398.65, 400.83, 404.24, 406.55, 408.52], dtype=np.float32)
    assert x.shape == y.shape == (60,)
    return {"x": x, "y": y}


# --------------------------------------------------------------------------- #
# German credit
# --------------------------------------------------------------------------- #
def load_german_credit(data_dir=None):
    """German credit design matrix, as the reference's model body builds it.

    ``models.py:860-881``: whitespace separated ``german.data`` (21 columns);
    object columns -> integer codes in ``np.unique`` order; numeric columns ->
    ``(c - mean) / std`` with pandas' ``std`` (ddof=1); an intercept column of
    ones is prepended to the numerics; ``status = (col 20 == 1)``.
    ``models.py:889-892``: ``all_x = concat([numericals] + one_hots, 1)`` in
    float32.  Returns ``X`` [N, F] float32 and ``y`` [N] float32 in {0, 1}.
    """
    import pandas as pd

    path = os.path.join(_data_dir(data_dir), "german.data")
    data = pd.read_csv(path, sep=r"\s+", header=None)
    n = len(data)
    numericals = [np.ones([n])]
    categoricals = []
    for col in data.columns[:-1]:
        column = data[col]
        if not pd.api.types.is_numeric_dtype(column):
            vals = column.to_numpy().astype(str)
            levels = {u: i for i, u in enumerate(np.unique(vals))}
            categoricals.append(np.array([levels[v] for v in vals]))
        else:
            column = column.astype(np.float64)
            numericals.append(((column - column.mean()) / column.std()).to_numpy())
    numericals = np.array(numericals).T.astype(np.float32)
    blocks = [numericals]
    for c in categoricals:
        depth = int(c.max()) + 1
        blocks.append(np.eye(depth, dtype=np.float32)[c])
    X = np.concatenate(blocks, axis=1).astype(np.float32)
    y = np.array(data[20] == 1, dtype=np.float32)
    return {"X": np.ascontiguousarray(X), "y": y}


def synthetic_german_credit(n=1000, f=25, seed=20190603):
    """Synthetic German-credit-shaped problem (BASELINE configs[1]).

    X = [1 | N(0,1)^{n x (f-1)}] in float32, beta* ~ N(0,1)/sqrt(f),
    y ~ Bernoulli(sigmoid(X beta*)).
    """
    rng = np.random.default_rng(seed)
    X = np.concatenate([np.ones((n, 1)), rng.standard_normal((n, f - 1))], axis=1)
    beta = rng.standard_normal(f) / np.sqrt(f)
    p = 1.0 / (1.0 + np.exp(-X @ beta))
    y = (rng.random(n) < p).astype(np.float32)
    return {"X": np.ascontiguousarray(X.astype(np.float32)), "y": y}


# --------------------------------------------------------------------------- #
# Radon
# --------------------------------------------------------------------------- #
def load_radon(state_code, data_dir=None):
    """Radon data for one state: restates reference ``models.py:706-760``.

    Returns county index ``c`` [N] int32 (0-based), log-uranium ``u`` [J]
    float32, floor indicator ``x`` [N] float32 and log-radon ``y`` [N] float32.
    """
    import pandas as pd

    d = _data_dir(data_dir)
    srrs2 = pd.read_csv(os.path.join(d, "srrs2.dat"))
    srrs2.columns = srrs2.columns.map(str.strip)
    srrs_mn = srrs2.assign(fips=srrs2.stfips * 1000 + srrs2.cntyfips)[srrs2.state == state_code]

    cty = pd.read_csv(os.path.join(d, "cty.dat"))
    cty_mn = cty[cty.st == state_code].copy()
    cty_mn["fips"] = 1000 * cty_mn.stfips + cty_mn.ctfips

    srrs_mn = srrs_mn.assign(county=srrs_mn.county.str.strip())

    counties = srrs_mn[["county", "fips"]].drop_duplicates()
    county_map_uranium = {a: b for a, b in zip(counties["county"], range(len(counties["county"])))}
    # label-indexed below exactly as the reference does (RangeIndex of the merge)
    uranium_levels = cty_mn.merge(counties, on="fips")["Uppm"]

    srrs_mn_new = srrs_mn.merge(cty_mn[["fips", "Uppm"]], on="fips")
    srrs_mn_new = srrs_mn_new.drop_duplicates(subset="idnum")
    srrs_mn_new = srrs_mn_new.assign(county=srrs_mn_new.county.str.strip())
    mn_counties = srrs_mn_new.county.unique()
    county_lookup = dict(zip(mn_counties, range(len(mn_counties))))

    county = srrs_mn_new.county.map(county_lookup).to_numpy().astype(np.int32)
    radon = srrs_mn_new.activity.to_numpy().astype(np.float64)
    log_radon = np.log(radon + 0.1)
    floor_measure = srrs_mn_new.floor.to_numpy()

    n_county = srrs_mn_new.groupby("county")["idnum"].count()
    uranium = np.zeros(len(n_county), dtype=np.float32)
    for k in county_lookup:
        uranium[county_lookup[k]] = uranium_levels[county_map_uranium[k]]
    uranium = [(np.log(ur) if ur > 0.0 else 0.0) for ur in uranium]

    return {
        "county": county,
        "u": np.float32(uranium),
        "x": np.float32(floor_measure),
        "y": np.float32(log_radon).reshape(-1),
    }


def synthetic_radon(n=1_000_000, j=10_000, seed=20190603):
    """Scaled synthetic radon: county sizes ~ Multinomial(n; Dirichlet(1)),
    sorted by county; x ~ Bern(0.17); u ~ N(0,1); y from the model with
    (mua, b1, b2) = (1.2, 0.7, -0.6)."""
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.ones(j))
    sizes = rng.multinomial(n, w)
    county = np.repeat(np.arange(j, dtype=np.int32), sizes)
    u = rng.standard_normal(j).astype(np.float32)
    x = (rng.random(n) < 0.17).astype(np.float32)
    m = 1.2 + 0.7 * u + rng.standard_normal(j)
    y = (m[county] + x * (-0.6) + rng.standard_normal(n)).astype(np.float32)
    return {"county": county, "u": u, "x": x, "y": y}


# --------------------------------------------------------------------------- #
# Election / electric (python-literal data modules)
# --------------------------------------------------------------------------- #
def load_election(data_dir=None):
    """``data/election88.py`` fields used by reference ``models.py:984-989``.

    ``state`` is kept 1-based exactly as stored; the one-hot out-of-range rule
     is applied where the index is consumed.
    """
    mod = runpy.run_path(os.path.join(_data_dir(data_dir), "election88.py"))
    d = mod["data"]
    return {
        "n_state": int(d["n_state"]),
        "black": np.asarray(d["black"], dtype=np.float32),
        "female": np.asarray(d["female"], dtype=np.float32),
        "state": np.asarray(d["state"], dtype=np.int32),
        "y": np.asarray(d["y"], dtype=np.float32),
    }


def load_electric(data_dir=None):
    """``data/electric.py`` fields used by reference ``models.py:1037-1045``."""
    mod = runpy.run_path(os.path.join(_data_dir(data_dir), "electric.py"))
    d = mod["data"]
    return {
        "n_pair": int(d["n_pair"]),
        "n_grade": int(d["n_grade"]),
        "n_grade_pair": int(d["n_grade_pair"]),
        "grade": np.asarray(d["grade"], dtype=np.int32),
        "grade_pair": np.asarray(d["grade_pair"], dtype=np.int32),
        "pair": np.asarray(d["pair"], dtype=np.int32),
        "treatment": np.asarray(d["treatment"], dtype=np.float32),
        "y": np.asarray(d["y"], dtype=np.float32),
    }


def onehot_index(idx, depth):
    """``tf.one_hot(idx, depth)`` semantics as an index map: in-range indices
    are kept, anything outside [0, depth) selects nothing (-1).  This is what
    the reference's 1-based Stan indices hit (``models.py:978,1016-1018``)."""
    idx = np.asarray(idx, dtype=np.int64)
    return np.where((idx >= 0) & (idx < depth), idx, -1).astype(np.int32)
