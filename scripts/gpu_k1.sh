mkdir -p gpurun_out
python -m pytest tests/test_gpu_mh_mvn.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -6
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_k1.json 2> gpurun_out/bench_k1.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_k1.json').read().strip().splitlines()[-1])
print('K1 kernel_ms', d['roofline']['kernel_ms'], 'value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'hbm frac', d['roofline']['frac'], d['quality'])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --accept reference > gpurun_out/bench_k1ref.json 2>> gpurun_out/bench_k1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_k1ref.json').read().strip().splitlines()[-1])
print('K1(ref accept) kernel_ms', d['roofline']['kernel_ms'], 'value %.3e'%d['value'])
PY
