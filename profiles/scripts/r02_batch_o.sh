#!/bin/bash
# round 2, GPU batch O: time_series with on-chip one-thread-per-chain state, new tests, full suite, radon_synth stream bench (1 GPU, short)
mkdir -p gpurun_out
timeout 600 python profiles/prof_simt.py time_series,8schools 2>&1 | grep -E "C +(4096|16384|131072|1048576)"
ARP_HMC_ONCHIP=0 timeout 600 python profiles/prof_simt.py time_series 2>&1 | grep -E "C +(16384|131072)" | sed 's/^/onchip off: /'
timeout 2000 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 900 python bench.py --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --steps 2 --warmup 1 > gpurun_out/r02o_bench_radon_synth.json 2> gpurun_out/r02o_bench_radon_synth.err
tail -c 2500 gpurun_out/r02o_bench_radon_synth.json; tail -5 gpurun_out/r02o_bench_radon_synth.err
