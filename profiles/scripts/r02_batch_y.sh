#!/bin/bash
# round 2, GPU batch Y: a quarter of the epilogue's exponentials on the FMA pipe (packed degree-5 polynomial) vs all on MUFU
mkdir -p gpurun_out
for m in pk7 pk23 pk7 pk23; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02y_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f accept %.4f' % (d['value'], d['ms_per_step'], d['ess']['acceptance_rate']))"
done
ARP_LIB_F32=build_dev/libarp_pk23.so timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "german_synth or single_leapfrog or confident or log_likelihood or internal_momenta or fp16_overflow" 2>&1 | grep -E "passed|failed|FAILED" | head
ARP_LIB_F32=build_dev/libarp_pk23.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "tcgen05 " | cut -c1-230
