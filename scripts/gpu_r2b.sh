# round 2, call B: K1 whitened-decision kernel -- parity tests + c2 bench
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_mh_mvn.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -15 ) 2>&1
python bench.py --workload c2 --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo rc=$?
tail -3 gpurun_out/r2b_bench_c2.err; cut -c1-1800 gpurun_out/r2b_bench_c2.json
python bench.py --workload c2 --no-secondary --no-cpu-baseline --no-e2e --variant 2 --steps 5 | cut -c1-200
python scripts/bench_k1_dims.py 2>&1 | tail -12
