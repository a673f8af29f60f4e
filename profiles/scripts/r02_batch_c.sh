#!/bin/bash
# round 2, GPU batch C: full parity suite (new VI / tuning-grid tests), typical-set precision, same-box A/B against the
# round-1 kernel, ncu of the ESS FFT kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc.py::test_tc_gradient_elementwise 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|^tests/" | head -120 > gpurun_out/r02c_pytest.log
timeout 600 python profiles/diag/diag_tc_typical.py 25 > gpurun_out/r02c_typical_f25.log 2>&1
timeout 600 python profiles/diag/diag_tc_typical.py 62 > gpurun_out/r02c_typical_f62.log 2>&1
timeout 600 python profiles/diag/diag_precision.py german > gpurun_out/r02c_diag_precision.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
ARP_LIB_F32=build_dev/libarp_r01.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02c_bench_r01lib.json 2> gpurun_out/r02c_bench_r01lib.err
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02c_bench2.json 2> gpurun_out/r02c_bench2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ess_fft -c 1 -o gpurun_out/r02c_ess_fft python profiles/prof_ess.py 835584 1000 1 > gpurun_out/r02c_ncu_ess.log 2>&1
tail -3 gpurun_out/r02c_pytest.log
cat gpurun_out/r02c_typical_f25.log
python - <<'PY'
import json
for n in ("", "_r01lib", "2"):
    try:
        d = json.loads(open("gpurun_out/r02c_bench%s.json" % n).read().strip().splitlines()[-1])
        print(n or "cur", "value %.4g e2e %.4g ms %.2f acc %.4f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["ess"]["acceptance_rate"], d["clocks"]))
    except Exception as e:
        print(n, "failed", e)
PY
