for v in default b8 default2; do
  if [ $v = default ] || [ $v = default2 ]; then unset FLOODSR_B200_LIB; else export FLOODSR_B200_LIB=$PWD/build/variants/lib_$v.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/abv_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/abv_$v.json')); m=d['modes']['fp16']; print('$v', round(d['ms_per_step'],2), d['stage_ms_per_step_rank0']['head'], d['stage_ms_per_step_rank0']['lr_conv'], '| fp16', round(m['ms_per_step'],2), m['stage_ms_per_step_rank0']['head'], m['stage_ms_per_step_rank0']['lr_conv'], d['clocks']['sm_mhz'])"
done
