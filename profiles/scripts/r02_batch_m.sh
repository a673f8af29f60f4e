#!/bin/bash
# round 2, GPU batch M: bisect continued (round-2 commits without the bias slot and with the fast exp), VI cluster kernel tests
mkdir -p gpurun_out
for m in r01 nb8c033be nb098ec34 nb2989168 r01 nb8c033be nb098ec34 nb2989168; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02m_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
timeout 1200 python -m pytest tests/test_gpu_ess_vi.py tests/test_gpu_reference_golden.py -m gpu -q 2>&1 | tail -15
for m in election electric; do
timeout 600 python bench.py --model $m --inference VI --method dVIP --steps 2 --warmup 1 > gpurun_out/r02m_vi_$m.json 2> gpurun_out/r02m_vi_$m.err; tail -c 1500 gpurun_out/r02m_vi_$m.json; tail -3 gpurun_out/r02m_vi_$m.err
done
