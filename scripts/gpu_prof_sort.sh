mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_post.py -m gpu -q -x 2>&1 | tail -5
python scripts/prof_sort.py > gpurun_out/prof_sort_plain.log 2>&1 || exit 1
python scripts/bench_post.py > gpurun_out/bench_post.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rs_scatter_kernel|rs_hist_kernel" --launch-skip 2 -c 2 -o gpurun_out/prof_sort_full -f python scripts/prof_sort.py > gpurun_out/ncu_sort.log 2>&1
tail -3 gpurun_out/ncu_sort.log
