mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_n2.err; echo rc=$?; cat gpurun_out/bench_ref_n2.json | cut -c1-600
