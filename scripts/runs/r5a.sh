# round-2 late session: LU launch list + one full capture of the cluster panel kernel on the ONERA M6 system (N = 7376),
# and the A/B of the L2-resident slice of A in the GMRES matvec
mkdir -p gpurun_out/r5a
timeout 120 python scripts/profile_case.py onera_m6 --solver LU --repeat 3 > gpurun_out/r5a/lu_time.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r5a/lu_launches.csv python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5a/lu_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lu_panel_coop -s 60 -c 1 -o gpurun_out/r5a/lu_panel_cl python scripts/profile_case.py onera_m6 --solver LU > gpurun_out/r5a/ncu_panel.log 2>&1
timeout 200 python scripts/gmres_ab.py onera_m6 - MACHLINE_GEMV_L2_PIN_MB=48 MACHLINE_GEMV_L2_PIN_MB=80 MACHLINE_GEMV_L2_PIN_MB=104 > gpurun_out/r5a/ab_m6.log 2>&1
cat gpurun_out/r5a/lu_time.log | tail -5
cat gpurun_out/r5a/ab_m6.log
ls -la gpurun_out/r5a
