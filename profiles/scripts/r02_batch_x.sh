#!/bin/bash
# round 2, GPU batch X (final build): full GPU suite, smoke(), default bench line with CPU baseline, reference arm,
# ncu --set full + source page of the tcgen05 HMC kernel with the packed epilogue, launch list of the bench command
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 2500 gpurun_out/r02_bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_german_tcs_hmc -s 1 -c 1 -f -o /tmp/r02_tcs python profiles/prof_hmc.py 0 16384 40 > gpurun_out/r02_ncu_tcs.log 2>&1
ncu -i /tmp/r02_tcs.ncu-rep --page raw --csv > gpurun_out/r02_tcs_raw.csv 2>/dev/null
ncu -i /tmp/r02_tcs.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_tcs_source_page.csv.gz
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/r02_launches_bench.log 2>&1
ls -la gpurun_out | tail -8
