# panel v2.4 (keys pushed before the parking, owner CTA by ballot, staged U-row loads): tests, hashes, stamps, ncu capture, N = 28k
mkdir -p gpurun_out/r5h
timeout 300 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_lu.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5h/pytest_lu.log 2>&1
tail -3 gpurun_out/r5h/pytest_lu.log
for n in 1000 4000; do timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5h/ab.log 2>&1; done
for n in 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5h/ab.log 2>&1
  timeout 120 python scripts/lu_ab.py $n >> gpurun_out/r5h/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 2000 >> gpurun_out/r5h/ab.log 2>&1
MACHLINE_LU_PANEL_DBG=1 timeout 120 python scripts/lu_ab.py 7376 --reps 1 >> gpurun_out/r5h/ab.log 2>&1
cat gpurun_out/r5h/ab.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lu_panel_cl2 -s 60 -c 1 -o gpurun_out/r5h/lu_panel_cl2 python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5h/ncu_panel.log 2>&1
timeout 250 python scripts/lu_ab.py 28403 --reps 2 > gpurun_out/r5h/lu28k.log 2>&1; cat gpurun_out/r5h/lu28k.log
