#!/bin/bash
# round 2, GPU batch Z: radon gradient sweep fused with the kicks: tests, radon_synth / radon PA benches; then batch Y (FMA-pipe exponentials)
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in 16 4; do
timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline --stream_window $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('radon_synth W$w value %.4g ms %.1f frac %.3f accept %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['ess']['acceptance_rate']))"
done
timeout 600 python profiles/prof_simt.py radon,radon_stddvs 2>&1 | grep -E "C +(16384|131072) " | grep "lpc  8"
bash profiles/scripts/r02_batch_y.sh
