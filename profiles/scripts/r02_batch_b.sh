#!/bin/bash
# round 2, GPU batch B: failing tests with full output, ESS (fused), exp accuracy modes A/B (timing + precision)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tc.py::test_tc_gradient_elementwise 2>&1 | tail -60 > gpurun_out/r02b_pytest.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "elementwise or many_chains" 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -80 > gpurun_out/r02b_pytest_tc.log
timeout 600 python profiles/prof_ess.py > gpurun_out/r02b_ess.log 2>&1
for m in exp0 exp1 exp2; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02b_bench_$m.json 2> gpurun_out/r02b_bench_$m.err
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python profiles/diag/diag_precision.py 2>&1 | grep german > gpurun_out/r02b_diag_$m.log
done
tail -5 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_ess.log
python - <<'PY'
import json
for n in ("exp0", "exp1", "exp2"):
    try:
        d = json.loads(open("gpurun_out/r02b_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g ms %.2f acc %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["ess"]["acceptance_rate"]))
    except Exception as e:
        print(n, "failed", e)
PY
