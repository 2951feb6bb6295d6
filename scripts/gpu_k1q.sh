timeout 600 python -m pytest tests/test_gpu_mh_mvn.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -2
python bench.py --no-secondary --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('log', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --no-secondary --no-cpu-baseline --steps 20 --accept reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ref', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --no-secondary --no-cpu-baseline --steps 20 --thin 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('thin10', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
