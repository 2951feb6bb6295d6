# A/B: K1 ring depth 3 (default build) vs 4 (prebuilt variant)
mkdir -p gpurun_out
run() { python bench.py --workload c2 --no-secondary --no-cpu-baseline --no-e2e --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel_ms'])"; }
echo "NSLOT=3"; run
cp probayes_b200/csrc/libpbx.so /tmp/libpbx_default.so
cp probayes_b200/csrc/_variants/libpbx_ns4.so probayes_b200/csrc/libpbx.so
echo "NSLOT=4"; run
cp /tmp/libpbx_default.so probayes_b200/csrc/libpbx.so
