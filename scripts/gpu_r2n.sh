mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gibbs_mvn" -s 2 -c 1 -f -o gpurun_out/prof_k5_r2 python scripts/prof_gibbs.py > gpurun_out/ncu_k5.log 2>&1
tail -1 gpurun_out/ncu_k5.log
ncu -i gpurun_out/prof_k5_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k5.csv
ncu -i gpurun_out/prof_k5_r2.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_k5.csv
