PBX_NVCC_EXTRA="-DWS_NPROD=18 -DWS_THREADS=768" python -m probayes_b200.build --force > /dev/null 2>&1
cuobjdump -res-usage probayes_b200/csrc/libpbx.so 2>/dev/null | grep -A1 "mh_mvn_ws_kernelILi2ELb0ELb1ELb1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
timeout 300 python -m pytest tests/test_gpu_mh_mvn.py -m gpu -q -x 2>&1 | tail -2
for t in 1 10; do python bench.py --no-secondary --no-cpu-baseline --steps 30 --thin $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nprod18 thin$t', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done
