# ncu capture of the K1 whitened-decision kernel (one launch of the C2 walk)
mkdir -p gpurun_out
TAG=${1:-r2b}
python scripts/prof_kernels.py k1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mh_mvn_wd -s 1 -c 1 -o gpurun_out/prof_k1_$TAG -f python scripts/prof_kernels.py k1 > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
ncu -i gpurun_out/prof_k1_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_k1.csv
ncu -i gpurun_out/prof_k1_$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_source_k1.csv
ls -la gpurun_out | tail -5
