mkdir -p gpurun_out
python scripts/time_k4.py > gpurun_out/k4_times.log 2>&1; cat gpurun_out/k4_times.log
ncu --set full --clock-control none --import-source on -k regex:"grid_max_sumexp|grid_posterior2" -s 6 -c 2 -o gpurun_out/prof_k4_r2 -f python scripts/time_k4.py > gpurun_out/ncu_k4.log 2>&1
tail -2 gpurun_out/ncu_k4.log
ncu -i gpurun_out/prof_k4_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k4.csv
ncu -i gpurun_out/prof_k4_r2.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_k4.csv
python -m pytest tests/test_gpu_grid.py -m gpu -q 2>&1 | tail -3
