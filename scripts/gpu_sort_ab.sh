mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_post.py -m gpu -q -x 2>&1 | tail -3
python scripts/bench_post.py > gpurun_out/bench_post_minb3.log 2>&1
grep -A2 "argsort_normal_n16777216\|argsort_normal_n67108864" gpurun_out/bench_post_minb3.log
PBX_NVCC_EXTRA="-DRS_MINB=4" python -m probayes_b200.build --force > /dev/null 2>&1
python scripts/bench_post.py > gpurun_out/bench_post_minb4.log 2>&1
grep -A2 "argsort_normal_n16777216\|argsort_normal_n67108864" gpurun_out/bench_post_minb4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort.csv python scripts/prof_sort.py > gpurun_out/ncu_sort.log 2>&1
