#!/bin/bash
# Round-2 profiling pass on the B200 box (run under gpurun): launch lists + full captures of the tensor-core kernels.
# LAUNCHES = kernel launches per bench step (37 in both modes since the folded upsampling; 39 before).
# Every ncu command follows a plain run of the same command line that exited 0 (B200_PROFILING.md).  gpurun_out/ may hold at
# most 64 MiB: the 27-kernel convolution captures are reduced to their raw-page CSV on the box.
mkdir -p gpurun_out
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__sass_inst_executed_op_utcmma.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum"
for prec in ${PRECS:-fp32 fp16}; do
  B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes --precision $prec"
  $B > gpurun_out/plain_$prec.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * ${LAUNCHES:-37})) -c ${LAUNCHES:-37} --csv --log-file gpurun_out/launches_r2_$prec.csv $B > gpurun_out/ncu_launch_$prec.log 2>&1
  echo "launch list $prec rc=$?"
  $B > gpurun_out/plain_$prec.log 2>&1 &&
  ncu --set full --metrics $TM --clock-control none --import-source on -k regex:fused_hr -s 3 -c 1 -f -o gpurun_out/prof_fused_$prec $B > gpurun_out/ncu_fused_$prec.log 2>&1
  echo "fused $prec rc=$?"
  $B > gpurun_out/plain_$prec.log 2>&1 &&
  ncu --set full --metrics $TM --clock-control none -k regex:"conv_rows_tc|conv_tc_kernel" -s 81 -c 27 -f -o /tmp/prof_conv_$prec $B > gpurun_out/ncu_conv_$prec.log 2>&1
  echo "conv $prec rc=$?"
  ncu -i /tmp/prof_conv_$prec.ncu-rep --page raw --csv > gpurun_out/prof_conv_${prec}_raw.csv 2>/dev/null
  rm -f /tmp/prof_conv_$prec.ncu-rep
done
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes --precision fp32"
$B > gpurun_out/plain_fp32.log 2>&1 &&
ncu --set full --clock-control none -k regex:"tile_normalize|blend" -s 6 -c 2 -f -o gpurun_out/prof_mem_r2 $B > gpurun_out/ncu_mem.log 2>&1
echo "mem rc=$?"
du -sh gpurun_out; ls -la gpurun_out
