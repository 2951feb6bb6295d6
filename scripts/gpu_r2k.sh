mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|Warning\|^$" | tail -25
python scripts/time_k4.py | tail -5
python bench.py --workload c2 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo rc=$?; tail -2 gpurun_out/r2k_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_c2.json').read().strip().splitlines()[-1])
print(d['value'])
for k,v in d['secondary'].items():
    if 'small' in k or 'normalise' in k or 'gibbs' in k: print(k, v)
PY
