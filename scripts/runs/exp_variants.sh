mkdir -p gpurun_out/r2e
for v in "" _cta2 _cta2nc _cta3nc; do
  lib=/root/repo/machline_b200/libmachline_gpu$v.so
  echo "== variant '$v'" >> gpurun_out/r2e/t.log
  MACHLINE_GPU_LIB=$lib python scripts/profile_case.py onera_m6 --repeat 5 --no-solve >> gpurun_out/r2e/t.log 2>&1
  MACHLINE_GPU_LIB=$lib python scripts/profile_step.py 96 52 --max-iter 3 >> gpurun_out/r2e/t.log 2>&1
  MACHLINE_GPU_LIB=$lib ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sass__inst_executed_shared_loads,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum --clock-control none -k regex:aic_assemble -c 1 --csv python scripts/profile_case.py onera_m6 --no-solve 2>&1 | grep -E "aic_assemble" | awk -F'","' '{print $(NF-2), $NF}' >> gpurun_out/r2e/t.log
done
