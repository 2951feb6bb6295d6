# round 2, final N = 1 call: parity tests, the four bench workloads, reference arm, launch list
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -4 ) 2>&1 | grep -v "^$"
for w in c2 c3 c4 c5; do
  ( time timeout 900 python bench.py --workload $w > gpurun_out/r2t_bench_$w.json 2> gpurun_out/r2t_bench_$w.err ) 2>&1 | grep real
  echo "== $w rc=$?"; tail -2 gpurun_out/r2t_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2t_bench_$w.json').read().strip().splitlines()[-1])
    d.pop('secondary',None)
    print(json.dumps(d)[:1800])
except Exception as e: print('no line', e)
PY
done
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2t_ref_c2.json 2> gpurun_out/r2t_ref_c2.err ) 2>&1 | grep real
tail -1 gpurun_out/r2t_ref_c2.json | cut -c1-600
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2t_plain.json 2> gpurun_out/r2t_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2t_ncu_launch.log 2>&1
tail -1 gpurun_out/r2t_ncu_launch.log | cut -c1-300
