N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print("binding", d.get("cpu_binding")); print('N=$N value %.4e ms/step %.4f e2e %.3e kernel_ms %.3f rhat %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['quality']['rhat'], d['clocks']))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>>gpurun_out/bench_n$N.err | cut -c1-200
