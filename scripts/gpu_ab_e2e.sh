#!/bin/bash
# A/B of the end-to-end pipeline on the B200 box (run under gpurun): one line per variant, "name:ENV=.. ENV=.."
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  [ "$envs" = "$v" ] && envs="FSR_AB=1"
  env $envs FSR_RASTER_TIMING=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/e2e_$name.json 2> gpurun_out/e2e_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/e2e_{name}.json"))
    m = d.get("modes", {}).get("fp16", {})
    print(f"E2E {name:12s} fp32 dev {d['ms_per_step']:.2f} e2e {d['e2e']['ms_per_step']:.2f} (ceil {d['e2e'].get('copy_ceiling_ms')}) head {d['stage_ms_per_step_rank0']['head']} | fp16 dev {m.get('ms_per_step', 0):.2f} e2e {m.get('e2e', {}).get('ms_per_step', 0):.2f} head {m.get('stage_ms_per_step_rank0', {}).get('head')} clk {d['clocks']['sm_mhz']}")
except Exception as exc:
    print(f"E2E {name:12s} FAILED {exc}")
PY
  grep fsr_run_raster gpurun_out/e2e_$name.err | sed -n "9p;\$p" | cut -c1-700
done
