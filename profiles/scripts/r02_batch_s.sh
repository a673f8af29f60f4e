#!/bin/bash
# round 2, GPU batch S: stated occupancy (launch bounds) for the SIMT kernels: sweep + radon_synth; full GPU suite;
# ncu: launch list of the bench command, --set full of the tcgen05 HMC kernel (with source page), the VI cluster kernel
# and the ESS FFT kernel (raw pages as csv; the reports stay on the box)
mkdir -p gpurun_out
timeout 900 python profiles/prof_simt.py > gpurun_out/r02s_simt.log 2>&1; grep -E "C +(4096|16384|131072|1048576) " gpurun_out/r02s_simt.log | grep -E "lpc +8|8schools.*lpc +1:|time_series"
for w in 16 4; do
timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 --no_cpu_baseline --stream_window $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('radon_synth W$w value %.4g ms %.1f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/r02_launches_bench.log 2>&1
grep -c "gpu__time_duration" gpurun_out/r02_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_german_tcs_hmc -s 1 -c 1 -f -o /tmp/r02_tcs python profiles/prof_hmc.py 0 16384 40 > gpurun_out/r02_ncu_tcs.log 2>&1
ncu -i /tmp/r02_tcs.ncu-rep --page raw --csv > gpurun_out/r02_tcs_raw.csv 2>/dev/null
ncu -i /tmp/r02_tcs.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_tcs_source_page.csv.gz
timeout 900 ncu --set full --clock-control none -k regex:k_vi -s 1 -c 1 -f -o /tmp/r02_vi python bench.py --model election --inference VI --method dVIP --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/r02_ncu_vi.log 2>&1
ncu -i /tmp/r02_vi.ncu-rep --page raw --csv > gpurun_out/r02_vi_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
