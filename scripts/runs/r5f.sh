# device post-processing tests; LU state after the reverts (panel v2.2, four interchange launches)
mkdir -p gpurun_out/r5f
timeout 300 python -m pytest tests/test_gpu_post.py -x -q > gpurun_out/r5f/pytest_post.log 2>&1
tail -15 gpurun_out/r5f/pytest_post.log
timeout 300 python -m pytest tests/test_gpu_solvers.py -m gpu -x -q -k "lu or LU" > gpurun_out/r5f/pytest_lu.log 2>&1
tail -3 gpurun_out/r5f/pytest_lu.log
for n in 7376 10513; do
  timeout 120 python scripts/lu_ab.py $n --heavy >> gpurun_out/r5f/ab.log 2>&1
  timeout 120 python scripts/lu_ab.py $n >> gpurun_out/r5f/ab.log 2>&1
done
timeout 120 python scripts/lu_ab.py 2000 >> gpurun_out/r5f/ab.log 2>&1
cat gpurun_out/r5f/ab.log
