# cluster panel v2.1 (REDUX arg-max, reciprocal before the barrier, threads sized to the rows, exact rows per CTA): tests, bitwise A/B, phase stamps
mkdir -p gpurun_out/r5c
timeout 400 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_lu.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5c/pytest.log 2>&1
tail -3 gpurun_out/r5c/pytest.log
for n in 1000 4000 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5c/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5c/ab.log 2>&1
MACHLINE_LU_PANEL_DBG=1 timeout 120 python scripts/lu_ab.py 7376 --reps 1 >> gpurun_out/r5c/ab.log 2>&1
MACHLINE_LU_PANEL_DBG=1 timeout 120 python scripts/lu_ab.py 2000 --reps 1 >> gpurun_out/r5c/ab.log 2>&1
MACHLINE_LU_LOOKAHEAD=0 timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5c/ab.log 2>&1
cat gpurun_out/r5c/ab.log
timeout 100 python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5c/m6_lu.log 2>&1; tail -1 gpurun_out/r5c/m6_lu.log
