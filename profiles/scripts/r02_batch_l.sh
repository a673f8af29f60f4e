#!/bin/bash
# round 2, GPU batch L: (1) bisect the 2.8 % slowdown of the tcgen05 HMC kernel over the round-2 commits (dev builds),
# (2) ncu --set full of the SIMT HMC kernel per model (raw page as csv; the reports themselves are too big to bring back)
mkdir -p gpurun_out
for m in r01 c8c033be c098ec34 c2989168 fa r01 c8c033be c098ec34 c2989168 fa; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02l_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
for spec in "8schools 1048576 1" "radon 131072 8" "radon_stddvs 131072 8" "election 131072 8" "electric 131072 8" "time_series 131072 1"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none -k regex:k_hmc_run -s 1 -c 1 -f -o /tmp/r02l_simt_$1 python profiles/prof_simt.py --ncu $1 $2 $3 > gpurun_out/r02l_simt_$1.log 2>&1
  ncu -i /tmp/r02l_simt_$1.ncu-rep --page raw --csv > gpurun_out/r02l_simt_$1_raw.csv 2>/dev/null
  tail -1 gpurun_out/r02l_simt_$1.log
done
ls -la gpurun_out | tail -20
