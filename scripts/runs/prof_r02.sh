set -x
mkdir -p gpurun_out/r02f
python bench.py > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f/bench_ref.json 2> gpurun_out/r02f/bench_ref.err
# launch list of one bench step (warm-up 1 + 1 step), no extra probes
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02f/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-lu-probe --no-extra-probes > gpurun_out/r02f/bench_under_ncu.log 2>&1
# full captures of the three kernels of the step
ncu --set full --clock-control none --import-source on -k regex:gemv_n_partial -s 300 -c 1 -o gpurun_out/r02f/gemv python scripts/profile_case.py onera_m6 > gpurun_out/r02f/ncu_gemv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:arnoldi_tail_kernel -s 300 -c 1 -o gpurun_out/r02f/tail python scripts/profile_case.py onera_m6 > gpurun_out/r02f/ncu_tail.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aic_assemble -s 1 -c 1 -o gpurun_out/r02f/aic_sub python scripts/profile_case.py onera_m6 --no-solve > gpurun_out/r02f/ncu_aic_sub.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aic_assemble -s 1 -c 1 -o gpurun_out/r02f/aic_sup python scripts/profile_case.py agard_b --no-solve > gpurun_out/r02f/ncu_aic_sup.log 2>&1
ls -la gpurun_out/r02f
