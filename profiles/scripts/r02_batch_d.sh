#!/bin/bash
# round 2, GPU batch D: HMC kernel A/B on one box (round-1 library, per-chunk G accumulation off / on), precision of the
# per-chunk variant at typical-set states, ESS packed-FP32x2 A/B
mkdir -p gpurun_out
for m in r01 gch0 gch1 r01 gch1; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r02d_bench_$m.json 2> gpurun_out/r02d_bench_$m.err
  python - "$m" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02d_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
ARP_LIB_F32=build_dev/libarp_gch1.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gch1 F62 value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
ARP_LIB_F32=build_dev/libarp_gch0.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --features 62 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gch0 F62 value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
ARP_LIB_F32=build_dev/libarp_gch1.so timeout 600 python profiles/diag/diag_tc_typical.py 25 2>&1 | grep -E "typical|simt |tcgen05 " > gpurun_out/r02d_typical_gch1_f25.log
ARP_LIB_F32=build_dev/libarp_gch1.so timeout 600 python profiles/diag/diag_tc_typical.py 62 2>&1 | grep -E "typical|simt |tcgen05 " > gpurun_out/r02d_typical_gch1_f62.log
cat gpurun_out/r02d_typical_gch1_f25.log gpurun_out/r02d_typical_gch1_f62.log
ARP_LIB_F32=build_dev/libarp_gch1.so timeout 600 python profiles/prof_ess.py 2>&1 | sed 's/^/packed: /'
ARP_LIB_F32=build_dev/libarp_fftnp.so timeout 600 python profiles/prof_ess.py 2>&1 | sed 's/^/scalar: /'
ARP_LIB_F32=build_dev/libarp_gch1.so timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ess_vi.py -m gpu -q -k "not elementwise and not vi_ and not param_adjoints" 2>&1 | tail -5
