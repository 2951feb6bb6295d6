mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --no-secondary > gpurun_out/bench_ns.json 2> gpurun_out/bench_ns.err; tail -c 300 gpurun_out/bench_ns.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_ns.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
