python -m pytest tests/test_gpu_grid.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from probayes_b200.engine import get_engine
eng=get_engine(0)
M=S=4096
lj=torch.randn(M,S,dtype=torch.float64,device='cuda')*30-5000
for name,fn in [('max',lambda: eng.grid_max(lj)),('cond',lambda: eng.grid_conditionalise(lj))]:
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0=torch.cuda.Event(enable_timing=True); t1=torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): r=fn()
    t1.record(); torch.cuda.synchronize(); print(name,'ms', t0.elapsed_time(t1)/10)
gm=eng.grid_max(lj); gs=eng.grid_sumexp(lj,gm)
for _ in range(3): eng.grid_posterior(lj,gm,gs)
ms=[]
for _ in range(5):
    eng.grid_posterior(lj,gm,gs); ms.append(eng.last_kernel_ms())
print('posterior kernel ms', np.median(ms), 'GB/s alg', 2*8*M*S/np.median(ms)/1e6)
ms=[]
for _ in range(5):
    eng.grid_sumexp(lj,gm); ms.append(eng.last_kernel_ms())
print('sumexp ms', np.median(ms)); ms=[]
for _ in range(5):
    eng.grid_max(lj); ms.append(eng.last_kernel_ms())
print('max ms', np.median(ms), 'GB/s', 8*M*S/np.median(ms)/1e6)
PY
