#!/bin/bash
# round 2, GPU batch H: full GPU suite (no -x)
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02h_pytest.log
cat gpurun_out/r02h_pytest.log
