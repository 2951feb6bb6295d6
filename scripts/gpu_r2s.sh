mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mvn_logpdf_mma64" -s 2 -c 1 -f -o gpurun_out/prof_dens_r2 python scripts/prof_dens.py > gpurun_out/ncu_dens.log 2>&1
tail -1 gpurun_out/ncu_dens.log
ncu -i gpurun_out/prof_dens_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_dens.csv
ncu -i gpurun_out/prof_dens_r2.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_dens.csv
