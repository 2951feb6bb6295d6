# round 2, after the tensor-core Gibbs kernel: C5 again at N = 1, 2, 4, 8 on one 8-GPU box
mkdir -p gpurun_out
P=29600
for w in c5; do
  for n in 1 2 4 8; do
    P=$((P+1))
    EXTRA="--steps 5 --warmup 3 --no-cpu-baseline"
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --workload $w --gpus 1 $EXTRA > gpurun_out/r2_scale_${w}_n$n.json 2> gpurun_out/r2_scale_${w}_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --workload $w --gpus $n $EXTRA > gpurun_out/r2_scale_${w}_n$n.json 2> gpurun_out/r2_scale_${w}_n$n.err
    fi
    echo "$w n=$n rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_${w}_n$n.json').read().strip().splitlines()[-1])
    print('  value %.4g %s  ms/step %.3f  e2e %.4g  frac %s' % (d['value'], d['unit'], d['ms_per_step'], (d['e2e'] or {}).get('value') or float('nan'), d['roofline'].get('frac')))
except Exception as e: print('  no line', e)
PY
  done
done
