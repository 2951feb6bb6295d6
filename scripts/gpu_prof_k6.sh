mkdir -p gpurun_out
python scripts/prof_sort.py > gpurun_out/prof_k6_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:"rs_scatter_kernel|rs_hist_kernel|cp_tile_sum_kernel|cp_tile_scan_kernel|ex_partial_kernel" --launch-skip 19 -c 19 -o gpurun_out/prof_k6_full -f python scripts/prof_sort.py > gpurun_out/ncu_k6.log 2>&1
tail -2 gpurun_out/ncu_k6.log; ls -la gpurun_out/prof_k6_full.ncu-rep
