#!/bin/bash
# round 2, GPU batch F: HMC kernel A/B (running G sum in TMEM): groups of 0 (off) / 1 / 2 / 4 chunks vs the round-1 library
mkdir -p gpurun_out
for m in r01 g0 g1 g2 g4 r01 g2; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02f_bench_$m.json 2> gpurun_out/r02f_bench_$m.err
  python - "$m" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02f_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02f_bench_%s.err" % sys.argv[1]).read()[-300:])
PY
done
for m in g2 g4; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "tcgen05 " | sed "s/^/$m /" | cut -c1-200
done
for m in g0 g2; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m F62 value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
ARP_LIB_F32=build_dev/libarp_g2.so timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "not elementwise and not many_chains" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -30
