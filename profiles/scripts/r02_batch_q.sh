#!/bin/bash
# round 2, GPU batch Q: lane-parallel time_series, single-offset buffer flip: tests, sweep, benches
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -30
timeout 900 python profiles/prof_simt.py time_series,electric,radon,8schools > gpurun_out/r02q_simt.log 2>&1; grep -E "C +(100|4096|16384|131072|1048576) " gpurun_out/r02q_simt.log
timeout 900 python bench.py --model time_series --chains 8192 --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('time_series 8192 value %.4g e2e %.4g ms %.1f accept %.3f rhat %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['ess']['acceptance_rate'], d['ess']['rhat_max']))"
timeout 600 python bench.py --model time_series --inference VI --method NCP --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('time_series VI value %.4g ms %.1f' % (d['value'], d['ms_per_step']))"
