mkdir -p gpurun_out
python -m pytest tests/test_gpu_mh_normreg.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_k2.json').read().strip().splitlines()[-1])
print('K1 kernel_ms', d['roofline']['kernel_ms'], 'value %.3e'%d['value'])
print('roofline_stream', d.get('roofline_stream'))
for k,v in d.get('secondary',{}).items(): print(k, json.dumps(v))
PY
