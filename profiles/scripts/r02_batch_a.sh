#!/bin/bash
# round 2, GPU batch A: parity suite, precision diagnostics, ESS A/B, HMC kernel A/B (bias slot / truncation split)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02a_pytest.log
timeout 600 python profiles/diag/diag_precision.py > gpurun_out/r02a_diag_precision.log 2>&1
timeout 600 python profiles/prof_ess.py > gpurun_out/r02a_ess.log 2>&1
ARP_LIB_F32=build_dev/libarp_minb1.so timeout 600 python profiles/prof_ess.py > gpurun_out/r02a_ess_minb1.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
ARP_LIB_F32=build_dev/libarp_nobias.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02a_bench_nobias.json 2> gpurun_out/r02a_bench_nobias.err
ARP_LIB_F32=build_dev/libarp_truncsplit.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02a_bench_truncsplit.json 2> gpurun_out/r02a_bench_truncsplit.err
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 > gpurun_out/r02a_bench_f62.json 2> gpurun_out/r02a_bench_f62.err
tail -3 gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_ess.log
python - <<'PY'
import json
for n in ("", "_nobias", "_truncsplit", "_f62"):
    try:
        d = json.loads(open("gpurun_out/r02a_bench%s.json" % n).read().strip().splitlines()[-1])
        print(n or "bias", "value %.4g e2e %.4g ms %.2f acc %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["ess"]["acceptance_rate"]))
    except Exception as e:
        print(n, "failed", e)
PY
