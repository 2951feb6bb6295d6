set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
python bench.py --steps 5 --warmup 3 --accept reference --no-cpu-baseline > gpurun_out/bench2_ref.json 2>> gpurun_out/bench2.err; cat gpurun_out/bench2_ref.json
python bench.py --steps 5 --warmup 3 --variant 1 --no-cpu-baseline > gpurun_out/bench2_v1.json 2>> gpurun_out/bench2.err; cat gpurun_out/bench2_v1.json
python bench.py --steps 3 --warmup 3 --chains 65536 --walk-steps 1000 --no-cpu-baseline > gpurun_out/bench2_64k.json 2>> gpurun_out/bench2.err; cat gpurun_out/bench2_64k.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mh_mvn_ws -s 1 -c 1 -o gpurun_out/prof_k1_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2b.log 2>&1
ls -la gpurun_out
