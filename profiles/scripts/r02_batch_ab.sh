#!/bin/bash
# round 2, GPU batch AB: time_series lane-parallel evaluation in mixed precision (level in double, the rest in fp32)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "time_series or lanes or vi_steps or learnable_b or param_adjoints or streaming or interleaved" 2>&1 | tail -6
timeout 600 python profiles/prof_simt.py time_series 2>&1 | grep -E "C +(100|4096|16384|131072) "
