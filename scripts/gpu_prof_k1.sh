mkdir -p gpurun_out
python -m pytest tests/test_gpu_mh_mvn.py -m gpu -q -x 2>&1 | tail -3
python scripts/prof_kernels.py k1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mh_mvn_ws -s 1 -c 1 -o gpurun_out/prof_k1_r1i -f python scripts/prof_kernels.py k1 > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
