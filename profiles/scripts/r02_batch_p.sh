#!/bin/bash
# round 2, GPU batch P: full GPU suite (fused sweeps, time_series in double, streaming CLI test), SIMT sweep, radon_synth
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -40
timeout 900 python profiles/prof_simt.py > gpurun_out/r02p_simt.log 2>&1; grep -E "C +(16384|131072|1048576)" gpurun_out/r02p_simt.log
timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/r02p_bench_radon_synth.json 2> gpurun_out/r02p_bench_radon_synth.err
python -c "
import json; d=json.loads(open('gpurun_out/r02p_bench_radon_synth.json').read().strip().splitlines()[-1]); print('radon_synth W16 value %.4g e2e %.4g ms %.1f roofline %.3g GB/s frac %.3f accept %.3f rhat %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['ess']['acceptance_rate'], d['ess']['rhat_max']))"
timeout 900 python bench.py --model time_series --chains 8192 --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('time_series 8192 value %.4g e2e %.4g ms %.1f accept %.3f rhat %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['ess']['acceptance_rate'], d['ess']['rhat_max']))"
