#!/bin/bash
# Final N = 1 records of the round (run under gpurun): bench lines for the three configs, the e2e pipeline timeline, and the
# profiling pass of the fp32 mode (launch list, full captures of the fused and the convolution kernels).
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_weak_n1.json 2> gpurun_out/bench_weak_n1.err; echo "weak rc=$?"
python bench.py --config c3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "c3 rc=$?"
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; echo "c4 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; echo "reference rc=$?"
FSR_RASTER_TIMING=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/raster_timing_final.err
grep fsr_run_raster gpurun_out/raster_timing_final.err | sed -n '5p;$p' > gpurun_out/raster_timeline_final.txt
PRECS=fp32 bash scripts/gpu_profile_r2.sh > gpurun_out/profile_final.log 2>&1
tail -5 gpurun_out/profile_final.log
python - <<'PY'
import json
for n in ("weak", "c3", "c4"):
    d = json.load(open(f"gpurun_out/bench_{n}_n1.json"))
    m = d["modes"]["fp16"]
    print(n, "fp32", round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), "roof", round(d["roofline"]["frac"], 3),
          "| fp16", round(m["value"]), round(m["ms_per_step"], 2), "e2e", round(m["e2e"]["value"]), round(m["e2e"]["ms_per_step"], 2), d["clocks"])
    print("   stages", d["stage_ms_per_step_rank0"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
