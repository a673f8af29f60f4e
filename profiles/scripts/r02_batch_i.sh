#!/bin/bash
# round 2, GPU batch I: HMC kernel A/B: round-1 library, split G accumulators off (gs0) / on (gs1); precision at typical-set
# states; tensor-core test file on gs1
mkdir -p gpurun_out
for m in r01 gs0 gs1 r01 gs1 gs0; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02i_bench_$m.json 2> gpurun_out/r02i_bench_$m.err
  python - "$m" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02i_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02i_bench_%s.err" % sys.argv[1]).read()[-300:])
PY
done
for m in gs0 gs1; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "tcgen05 " | sed "s/^/$m F25 /" | cut -c1-230
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python profiles/diag/diag_tc_typical.py 62 2>&1 | grep -E "tcgen05 " | sed "s/^/$m F62 /" | cut -c1-230
done
for m in gs0 gs1; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m F62 value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
ARP_LIB_F32=build_dev/libarp_gs1.so timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED" | head -30
