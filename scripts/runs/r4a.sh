mkdir -p gpurun_out/r4a
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_gmres.py tests/test_gpu_solver_histories.py -m gpu -x -q > gpurun_out/r4a/pytest.log 2>&1
tail -3 gpurun_out/r4a/pytest.log
timeout 300 python scripts/gmres_ab.py onera_m6 MACHLINE_GMRES_TAIL=1,MACHLINE_GMRES_D2H_COPY=1 MACHLINE_GMRES_D2H_COPY=1 - MACHLINE_GEMV_L2_PIN_MB=32 MACHLINE_GEMV_L2_PIN_MB=48 MACHLINE_GEMV_L2_PIN_MB=64 MACHLINE_GEMV_L2_PIN_MB=80 MACHLINE_GEMV_L2_PIN_MB=96 MACHLINE_GEMV_L2_PIN_MB=112 MACHLINE_GMRES_SHARDED=1 MACHLINE_GMRES_SHARDED=1,MACHLINE_GMRES_D2H_COPY=1 MACHLINE_GMRES_SHARDED=1,MACHLINE_GEMV_L2_PIN_MB=64 > gpurun_out/r4a/ab_m6.log 2>&1
cat gpurun_out/r4a/ab_m6.log
timeout 300 python scripts/gmres_ab.py x --synthetic 1 MACHLINE_GMRES_TAIL=1,MACHLINE_GMRES_D2H_COPY=1 - MACHLINE_GEMV_L2_PIN_MB=48 MACHLINE_GEMV_L2_PIN_MB=64 MACHLINE_GEMV_L2_PIN_MB=80 MACHLINE_GEMV_L2_PIN_MB=96 > gpurun_out/r4a/ab_syn.log 2>&1
cat gpurun_out/r4a/ab_syn.log
