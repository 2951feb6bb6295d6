mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python scripts/prof_kernels.py k1 gibbs tiles > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
prof() {
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s 1 -c 1 \
      -o gpurun_out/prof_$1 python scripts/prof_kernels.py $3 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
prof k1_final mh_mvn_ws k1
prof gibbs_final gibbs_mvn_kernel gibbs
prof tiles_final nr_tiles tiles
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_l.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_final.csv
