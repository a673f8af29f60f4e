#!/bin/bash
# round 2, GPU batch AA: block-wise streaming lag sums: tests, radon_synth with W = 16 / 4 / 64
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in 16 4 64; do
timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline --stream_window $w 2>gpurun_out/r02aa_$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('radon_synth W$w value %.4g ms %.1f frac %.3f achieved %.4g accept %.3f rhat %.3f ess/1000 %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['ess']['acceptance_rate'], d['ess']['rhat_max'], d['ess']['ess_per_1000_grads_mean']))"
done
