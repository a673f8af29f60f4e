#!/bin/bash
# round 2: multi-GPU runs on ONE box with N GPUs (gpurun --gpus N -- bash profiles/scripts/r02_scale.sh N)
#   bench.py (BASELINE configs[1], weak and strong scaling), configs[4] (radon_synth + time_series, 8192 chains per GPU),
#   per-model table (bench_models.py)
N=${1:-8}
ONLY=${2:-all}     # "german": only the configs[1] lines; "configs4": only radon_synth + time_series (re-runs after kernel changes)
mkdir -p gpurun_out
run() {  # run NAME args...: bench.py on N ranks, JSON line -> gpurun_out/r02s_NAME_nN.json
  name=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r02s_${name}_n$N.json 2> gpurun_out/r02s_${name}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/r02s_${name}_n$N.json 2> gpurun_out/r02s_${name}_n$N.err
  fi
  python - "$name" "$N" <<'PY'
import json, sys
name, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/r02s_%s_n%s.json" % (name, n)).read().strip().splitlines()[-1])
    print("%s N=%s value %.4g e2e %.4g ms %.1f roofline %s %.3g frac %.3f accept %.3f rhat %.3f chains %d" % (
        name, n, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["bound"], d["roofline"]["achieved"],
        d["roofline"]["frac"], d["ess"]["acceptance_rate"], d["ess"]["rhat_max"], d["ess"]["chains_reduced"]))
except Exception as e:
    print(name, "failed", e, open("gpurun_out/r02s_%s_n%s.err" % (name, n)).read()[-800:])
PY
}
if [ "$ONLY" != configs4 -a "$ONLY" != time_series ]; then
run german_weak --steps 3 --warmup 3 --no_cpu_baseline
run german_strong --steps 3 --warmup 3 --no_cpu_baseline --scaling strong --chains 16384
fi
[ "$ONLY" = german ] && exit 0
[ "$ONLY" != time_series ] && run radon_synth --model radon_synth --chains 8192 --num_samples 100 --num_burnin_steps 100 --num_adaptation_steps 80 --stream_window 64 --steps 2 --warmup 1 --no_cpu_baseline
run time_series --model time_series --chains 8192 --steps 2 --warmup 1 --no_cpu_baseline
[ "$ONLY" = configs4 -o "$ONLY" = time_series ] && exit 0
if [ "$N" = 1 ]; then
  timeout 1500 python bench_models.py --out gpurun_out/r02_models_n$N.json > gpurun_out/r02_models_n$N.log 2>&1
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench_models.py --no_cpu_baseline --out gpurun_out/r02_models_n$N.json > gpurun_out/r02_models_n$N.log 2>&1
fi
grep -E "^\|" gpurun_out/r02_models_n$N.log | cut -c1-260
tail -3 gpurun_out/r02_models_n$N.log | cut -c1-300
