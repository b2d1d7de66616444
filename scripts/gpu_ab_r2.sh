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
run default FSR_DUMMY=1
run wait96 FSR_X3_WAIT_NS=96,96,96
run wait1000 FSR_X3_WAIT_NS=1000,400,1000
run wait2000 FSR_X3_WAIT_NS=2000,1000,2000
run chain24 FSR_X3_CHAIN=24
run chain96 FSR_X3_CHAIN=96
run nopair FSR_NO_CONV_PAIR=1
for c in 48 96; do
  FSR_X3_CHAIN=$c timeout 300 python tests/x3_error_budget.py --modes fp32 --out gpurun_out/x3_error_budget_chain$c.txt > /dev/null 2>&1
  grep "==" gpurun_out/x3_error_budget_chain$c.txt
  awk '{ if ($0 ~ /err/) print }' gpurun_out/x3_error_budget_chain$c.txt | sort -t'r' -k3 | tail -0
done
