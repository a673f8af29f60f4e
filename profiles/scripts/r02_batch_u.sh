#!/bin/bash
# round 2, GPU batch U: which packed fp32x2 instructions pay in the epilogue (mask: 1 adds, 2 multiplies, 4 head subtraction)
mkdir -p gpurun_out
for m in r01 pk0 pk7 pk1 pk2 pk4 pk3 pk5 pk6 pk7 pk0 pk3 pk5 pk6; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02u_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
