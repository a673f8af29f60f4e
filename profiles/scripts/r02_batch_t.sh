#!/bin/bash
# round 2, GPU batch T: packed fp32x2 epilogue (FADD2 / FMUL2 / FFMA2): micro-benchmark, kernel A/B, tensor-core tests
mkdir -p gpurun_out
./build_dev/epi | tee gpurun_out/r02t_epi.log
for m in r01 pk0 pk1 pk0 pk1; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02t_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
ARP_LIB_F32=build_dev/libarp_fpk1.so timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED" | head -20
ARP_LIB_F32=build_dev/libarp_fpk1.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "tcgen05 " | cut -c1-230
ARP_LIB_F32=build_dev/libarp_fpk1.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fpk1 value %.4g e2e %.4g ms %.2f accept %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['ess']['acceptance_rate']))"
