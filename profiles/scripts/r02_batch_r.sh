#!/bin/bash
# round 2, GPU batch R: buffer flip by swapping Vec objects (commit 9abba2e) vs by one offset (current): radon_synth (state
# in HBM), on-chip models
mkdir -p gpurun_out
for m in vecswap cur vecswap cur; do
  if [ $m = cur ]; then unset ARP_LIB_F32; else export ARP_LIB_F32=build_dev/libarp_$m.so; fi
  timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m radon_synth W16 value %.4g ms %.1f' % (d['value'], d['ms_per_step']))"
  timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline --stream_window 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m radon_synth W4 value %.4g ms %.1f' % (d['value'], d['ms_per_step']))"
done
unset ARP_LIB_F32
timeout 600 python -m pytest tests/test_gpu_ess_vi.py -m gpu -q -k "fft_oracle" 2>&1 | tail -3
