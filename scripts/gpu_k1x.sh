PBX_NVCC_EXTRA="-DPBX_K1_NOSTATS" python -m probayes_b200.build --force > /dev/null 2>&1
for t in 1 10; do python bench.py --no-secondary --no-cpu-baseline --steps 30 --thin $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nostats thin$t', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done
