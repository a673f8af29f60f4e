#!/bin/bash
# round 2, GPU batch W: reverse sweep specialised on the last leapfrog step (sl1), head by truncation in the packed epilogue (pk15)
mkdir -p gpurun_out
for m in pk7 sl1 pk15 pk7 sl1 pk15; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02w_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f accept %.4f' % (d['value'], d['ms_per_step'], d['ess']['acceptance_rate']))"
done
