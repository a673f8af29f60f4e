#!/bin/bash
# round 2, GPU batch K: per-kernel device times (ncu, short workload) of the round-1 library vs the current one
mkdir -p gpurun_out
for m in r01 fa; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02k_launches_$m.csv python profiles/prof_hmc.py 0 16384 40 > gpurun_out/r02k_$m.log 2>&1
grep -E "k_german|k_hmc" gpurun_out/r02k_launches_$m.csv | cut -d, -f5,13- | tail -12
done
# ncu --set full of the SIMT HMC kernel, one capture per model (second launch = the timed-size run)
for spec in "8schools 1048576 1" "radon 131072 8" "radon_stddvs 131072 8" "election 131072 8" "electric 131072 8" "time_series 131072 1"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hmc_run -s 1 -c 1 -f -o gpurun_out/r02k_simt_$1 python profiles/prof_simt.py --ncu $1 $2 $3 > gpurun_out/r02k_simt_$1.log 2>&1
  ncu -i gpurun_out/r02k_simt_$1.ncu-rep --page raw --csv > gpurun_out/r02k_simt_$1_raw.csv 2>/dev/null
  tail -1 gpurun_out/r02k_simt_$1.log
done
ls -la gpurun_out/*.ncu-rep | tail
