#!/bin/bash
# round 2, GPU batch J: full builds, exp mode 1 (fa) vs 0 (fe0) vs the round-1 library: bench + tensor-core tests
mkdir -p gpurun_out
for m in r01 fa fe0 fa fe0; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02j_bench_$m.json 2> gpurun_out/r02j_bench_$m.err
  python - "$m" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02j_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02j_bench_%s.err" % sys.argv[1]).read()[-300:])
PY
done
for m in fa fe0; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED" | head -30 | sed "s/^/$m /"
done
