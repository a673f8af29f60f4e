#!/bin/bash
# round 2, GPU batch N: tcgen05 kernel perturbations (truncation split, exp modes), full GPU suite with the new on-chip
# layouts, SIMT sweep, bench lines of the other BASELINE configs
mkdir -p gpurun_out
for m in r01 fa p1 p2 p3 p1 p2 p3; do
  ARP_LIB_F32=build_dev/libarp_$m.so timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline 2> gpurun_out/r02n_bench_$m.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"
done
timeout 2000 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 900 python profiles/prof_simt.py 8schools,radon,radon_stddvs,election,electric > gpurun_out/r02n_simt.log 2>&1; grep -E "C +(16384|131072|1048576)" gpurun_out/r02n_simt.log
for spec in "8schools CP 1048576" "radon NCP 16384" "election NCP 16384" "time_series NCP 16384"; do
  set -- $spec
  timeout 600 python bench.py --model $1 --method $2 --chains $3 --steps 3 --warmup 3 > gpurun_out/r02n_bench_$1.json 2> gpurun_out/r02n_bench_$1.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02n_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g e2e %.4g ms %.2f roofline %s %.3g/%.3g cpu %.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["bound"], d["roofline"]["achieved"], d["roofline"]["peak"], d.get("cpu_baseline", {}).get("value", 0)))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02n_bench_%s.err" % sys.argv[1]).read()[-600:])
PY
done
