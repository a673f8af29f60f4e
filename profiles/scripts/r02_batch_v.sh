#!/bin/bash
# round 2, GPU batch V: final build: full GPU suite, smoke(), N = 1 scale script
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash profiles/scripts/r02_scale.sh 1
