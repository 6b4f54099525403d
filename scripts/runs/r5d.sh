# LU: new TRSM (warp per 4 columns), fused row interchanges per pair, register/shuffle triangular solves + strip prefetch in the
# substitution, panel v2.2 (staged elimination loads): tests, hashes of x (must equal r5b / r5c), phase stamps, launch list
mkdir -p gpurun_out/r5d
timeout 400 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_sharded_lu.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5d/pytest.log 2>&1
tail -3 gpurun_out/r5d/pytest.log
for n in 1000 4000 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5d/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 2000 >> gpurun_out/r5d/ab.log 2>&1
timeout 120 python scripts/lu_ab.py 7376 >> gpurun_out/r5d/ab.log 2>&1
timeout 120 python scripts/lu_ab.py 10513 >> gpurun_out/r5d/ab.log 2>&1
MACHLINE_LU_PANEL_DBG=1 timeout 120 python scripts/lu_ab.py 7376 --reps 1 >> gpurun_out/r5d/ab.log 2>&1
cat gpurun_out/r5d/ab.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r5d/lu_launches.csv python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5d/lu_under_ncu.log 2>&1
timeout 100 python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5d/m6_lu.log 2>&1; tail -1 gpurun_out/r5d/m6_lu.log
