# K1 whitened-decision kernel: parity tests, c2 bench (device-timed only), ncu capture
mkdir -p gpurun_out
TAG=${1:-r2c}
python -m pytest tests/test_gpu_mh_mvn.py -m gpu -q -x 2>&1 | tail -4
python bench.py --workload c2 --no-secondary --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo rc=$?
tail -3 gpurun_out/${TAG}_bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['quality'])"
if [ "$2" != "noprof" ]; then bash scripts/gpu_prof_k1wd.sh $TAG | tail -3; fi
