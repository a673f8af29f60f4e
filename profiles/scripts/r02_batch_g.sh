#!/bin/bash
# round 2, GPU batch G: full GPU suite on the working tree (TCS_GCHUNK 0), bench line, per-model SIMT profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02g_pytest.log
cat gpurun_out/r02g_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
cat gpurun_out/r02g_bench.json
timeout 600 python profiles/prof_simt.py > gpurun_out/r02g_simt.log 2>&1
cat gpurun_out/r02g_simt.log | tail -40
