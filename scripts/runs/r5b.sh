# second-generation cluster panel: LU tests, bitwise A/B against the first cluster kernel, timings
mkdir -p gpurun_out/r5b
timeout 400 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_lu.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5b/pytest.log 2>&1
tail -4 gpurun_out/r5b/pytest.log
for n in 1000 4000 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5b/ab.log 2>&1
  MACHLINE_LU_PANEL_V1=1 timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5b/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5b/ab.log 2>&1
MACHLINE_LU_PANEL_V1=1 timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5b/ab.log 2>&1
cat gpurun_out/r5b/ab.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r5b/lu_launches.csv python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5b/lu_under_ncu.log 2>&1
tail -2 gpurun_out/r5b/lu_under_ncu.log
