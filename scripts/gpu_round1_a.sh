set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python __graft_entry__.py --smoke 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
python bench.py --steps 3 --warmup 3 --accept log --no-cpu-baseline > gpurun_out/bench1_log.json 2>> gpurun_out/bench1.err; cat gpurun_out/bench1_log.json
python bench.py --steps 3 --warmup 3 --chains 65536 --walk-steps 1000 --no-cpu-baseline > gpurun_out/bench1_64k.json 2>> gpurun_out/bench1.err; cat gpurun_out/bench1_64k.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mh_mvn -s 1 -c 1 -o gpurun_out/prof_k1_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
