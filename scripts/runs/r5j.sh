# final of the session: the bench line, then the whole GPU test suite
mkdir -p gpurun_out/r5j
timeout 120 python bench.py > gpurun_out/r5j/bench.json 2> gpurun_out/r5j/bench.err
head -c 300 gpurun_out/r5j/bench.json; echo
timeout 290 python -m pytest tests -m gpu -x -q > gpurun_out/r5j/pytest_gpu.log 2>&1
tail -4 gpurun_out/r5j/pytest_gpu.log
