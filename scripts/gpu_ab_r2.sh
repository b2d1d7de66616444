#!/bin/bash
# A/B measurements of the split-mode knobs on the B200 box (run under gpurun); prints one line per variant.
mkdir -p gpurun_out
run() {  # name, env assignments...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{name}.json"))
    print(f"AB {name:14s} {d['ms_per_step']:.3f} ms/step  stages {d['stage_ms_per_step_rank0']}  roof {d['roofline']['frac']:.3f}")
except Exception as exc:
    print(f"AB {name:14s} FAILED {exc}")
PY
}
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  run $name $envs
done
