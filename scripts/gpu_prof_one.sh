# usage: gpu_prof_one.sh <name> <kernel-regex> <prof_kernels.py arg>
mkdir -p gpurun_out
python scripts/prof_kernels.py $3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 1 -c 1 -o gpurun_out/prof_$1 python scripts/prof_kernels.py $3 > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
