# round 2, call A: GPU parity tests + the four bench workloads at N = 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) 2>&1
for w in c2 c3 c4 c5; do
  ( time timeout 900 python bench.py --workload $w > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err ) 2>&1 | grep real
  echo "== $w rc=$?"; tail -2 gpurun_out/r2a_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench_$w.json').read().strip().splitlines()[-1])
    d.pop('secondary',None)
    print(json.dumps(d)[:2500])
except Exception as e: print('no line', e)
PY
done
