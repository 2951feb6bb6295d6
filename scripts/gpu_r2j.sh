mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) 2>&1
python bench.py --workload c4 --no-cpu-baseline --steps 5 > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err; echo rc=$?; tail -2 gpurun_out/r2j_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['quality'], d['e2e'])
PY
python bench.py --workload c2 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err; echo rc=$?; tail -2 gpurun_out/r2j_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_c2.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline'])
print(d['secondary']['c4_normalise_marginals'])
print(d['secondary_n'])
PY
