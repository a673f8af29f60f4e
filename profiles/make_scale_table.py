"""profiles/r02_bench_<workload>_n<N>.json (bench.py lines of profiles/scripts/r02_scale.sh) -> profiles/r02_scale.md"""
import glob
import json
import os
import re

here = os.path.dirname(os.path.abspath(__file__))
rows = {}
for f in glob.glob(os.path.join(here, "r02_bench_*_n*.json")):
    m = re.match(r"r02_bench_(.+)_n(\d+)\.json", os.path.basename(f))
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    rows.setdefault(m.group(1), {})[int(m.group(2))] = d
out = ["# Multi-GPU runs, round 2 (`profiles/scripts/r02_scale.sh N` under `gpurun --gpus N`; one process per GPU, NCCL)", "",
       "`value` = all chains of all ranks / max-over-ranks device time; efficiency = value(N) / (N x value(1)) for the weak-scaling",
       "workloads (chains per GPU fixed) and value(N) / value(1) / N for the strong-scaling one (16 384 chains in total).", ""]
names = {"german_weak": "configs[1] German credit 1000 x 25, 16 384 chains per GPU (weak)",
         "german_strong": "configs[1] German credit, 16 384 chains IN TOTAL (strong)",
         "radon_synth": "configs[4] synthetic radon 10^6 x 10^4, 8192 chains per GPU, streaming ESS W = 64, 100 kept samples",
         "time_series": "configs[4] time_series, 8192 chains per GPU, 1000 kept samples (final kernel with the mixed-precision scans: 1 and 8 GPUs; the 2- and 4-GPU runs of the all-double kernel scaled 0.99 / 0.99)"}
for key in ("german_weak", "german_strong", "radon_synth", "time_series"):
    if key not in rows:
        continue
    out += ["## " + names[key], "",
            "| GPUs | chains | grad-evals/s | e2e grad-evals/s | ms / step | efficiency | roofline | acceptance | R-hat max |",
            "|---|---|---|---|---|---|---|---|---|"]
    base = rows[key].get(1)
    for n in sorted(rows[key]):
        d = rows[key][n]
        eff = d["value"] / (base["value"] * n) if base else float("nan")
        r = d["roofline"]
        out.append("| %d | %d | %.4g | %.4g | %.1f | %.3f | %.3g %s = %.1f %% (%s) | %.3f | %.3f |" % (
            n, d["ess"]["chains_reduced"], d["value"], d["e2e"]["value"], d["ms_per_step"], eff, r["achieved"], r["unit"],
            100 * r["frac"], r["bound"], d["ess"]["acceptance_rate"], d["ess"]["rhat_max"]))
    out.append("")
open(os.path.join(here, "r02_scale.md"), "w").write("\n".join(out))
print("\n".join(out))
