# panel v2.3 (candidate keys pulled after the barrier, REDUX + ballot arg-max), 128-thread row-interchange kernel
mkdir -p gpurun_out/r5e
timeout 400 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_lu.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5e/pytest.log 2>&1
tail -3 gpurun_out/r5e/pytest.log
for n in 1000 4000 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5e/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 2000 >> gpurun_out/r5e/ab.log 2>&1
timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5e/ab.log 2>&1
timeout 120 python scripts/lu_ab.py 10513 >> gpurun_out/r5e/ab.log 2>&1
MACHLINE_LU_PANEL_DBG=1 timeout 120 python scripts/lu_ab.py 7376 --reps 1 >> gpurun_out/r5e/ab.log 2>&1
cat gpurun_out/r5e/ab.log
