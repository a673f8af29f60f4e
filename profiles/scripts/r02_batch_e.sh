#!/bin/bash
# round 2, GPU batch E: HMC kernel A/B on one box: round-1 library vs the current kernel with the per-group G accumulation
# off (g0: exp mode 1, g0e0: exp mode 0) and on (g1l / g2l: groups of 1 / 2 chunks, late fetch; g2e: early fetch)
mkdir -p gpurun_out
for m in r01 g0 g0e0 g1l g2l g2e g1e r01 g2l; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02e_bench_$m.json 2> gpurun_out/r02e_bench_$m.err
  python - "$m" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02e_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
for m in g1l g2l; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "typical|tcgen05 " | sed "s/^/$m /"
done
for m in g0 g2l; do
ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m F62 value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
ARP_LIB_F32=build_dev/libarp_g2l.so timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "not elementwise" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -30
