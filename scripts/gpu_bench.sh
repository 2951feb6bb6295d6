mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; echo "rc=$?"
tail -3 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_latest.json').read().strip().splitlines()[-1])
sec=d.pop('secondary',{})
print(json.dumps(d, indent=1)[:3000])
for k,v in sec.items(): print(k, json.dumps(v))
PY
