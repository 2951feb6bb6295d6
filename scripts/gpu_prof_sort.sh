mkdir -p gpurun_out
timeout 120 python scripts/prof_sort.py > gpurun_out/prof_sort_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort.csv python scripts/prof_sort.py > gpurun_out/ncu_sort.log 2>&1
