mkdir -p gpurun_out
python scripts/prof_kernels.py tiles gibbs > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none -k regex:"nr_tiles|gibbs_mvn_kernel" -c 5 -f -o gpurun_out/prof_k2k5_r1i python scripts/prof_kernels.py tiles gibbs > gpurun_out/ncu_k2k5.log 2>&1
tail -1 gpurun_out/ncu_k2k5.log
