mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -m gpu -q -x 2>&1 | tail -40
timeout 300 python scripts/bench_post.py 2>&1 | tee gpurun_out/bench_post.log | tail -30
