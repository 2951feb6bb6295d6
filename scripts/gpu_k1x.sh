for u in 4 1; do
PBX_NVCC_EXTRA="-DWS_PUNROLL=$u" python -m probayes_b200.build --force > /dev/null 2>&1
cuobjdump -res-usage probayes_b200/csrc/libpbx.so 2>/dev/null | grep -A1 "mh_mvn_ws_kernelILi2ELb0ELb1ELb1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
python bench.py --no-secondary --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unroll$u', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
