# two GPUs: the sharded / multi-context tests after the LU changes; the LU hashes with the two-range interchange launches
mkdir -p gpurun_out/r5i
timeout 420 python -m pytest tests/test_gpu_sharded_lu.py tests/test_gpu_sharded_gmres.py tests/test_gpu_multi_ctx.py -m gpu -x -q > gpurun_out/r5i/pytest_2gpu.log 2>&1
tail -4 gpurun_out/r5i/pytest_2gpu.log
for n in 4000 7376; do timeout 120 python scripts/lu_ab.py $n --heavy --reps 2 >> gpurun_out/r5i/ab.log 2>&1; done
timeout 120 python scripts/lu_ab.py 7376 --reps 2 >> gpurun_out/r5i/ab.log 2>&1
cat gpurun_out/r5i/ab.log
