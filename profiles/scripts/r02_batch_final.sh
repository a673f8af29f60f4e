#!/bin/bash
# round 2, final GPU batch (final build): full GPU suite, smoke(), default bench line, per-model table at N = 1
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-400 gpurun_out/r02_bench_final.json
timeout 1500 python bench_models.py --out gpurun_out/r02_models_n1.json > gpurun_out/r02_models_n1.log 2>&1
grep -E "^\| (radon|time_series|8schools|election)[a-z_]* \| NCP" gpurun_out/r02_models_n1.log | cut -c1-260
