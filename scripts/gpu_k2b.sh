python -m pytest tests/test_gpu_mh_normreg.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from probayes_b200.engine import get_engine
eng=get_engine(0); rng=np.random.default_rng(2024)
lims=np.array([[-6.,6.],[-6.,6.],[0.001,10.]]); ex=np.array([[0,0],[0,0],[1,0]]); lg=np.zeros(3,int)
N,C=1_000_000,16384
x=eng.to_device(rng.normal(0,1,N)); y=eng.to_device(rng.normal(0,1,N))
th=eng.to_device(np.stack([rng.normal(-1,.001,C),rng.normal(1.5,.001,C),rng.uniform(.49,.51,C)]))
for _ in range(3): eng.normreg_logjoint(th,y,x,lims,ex,lg,variant=1)
ms=[]
for _ in range(7):
    eng.normreg_logjoint(th,y,x,lims,ex,lg,variant=1); ms.append(eng.last_kernel_ms())
m=np.median(ms); print('tiles C=16384 N=1e6 ms',m,'terms/s %.3e'%(N*C/m*1e3),'TFLOPs(5/term) %.2f'%(5*N*C/m*1e3/1e12), 'pipe instr/s %.3e'%(3*N*C/m*1e3))
for C2 in (2048, 4096, 65536):
    th2=eng.to_device(np.stack([rng.normal(-1,.001,C2),rng.normal(1.5,.001,C2),rng.uniform(.49,.51,C2)]))
    for _ in range(2): eng.normreg_logjoint(th2,y,x,lims,ex,lg,variant=1)
    ms=[]
    for _ in range(5):
        eng.normreg_logjoint(th2,y,x,lims,ex,lg,variant=1); ms.append(eng.last_kernel_ms())
    m=np.median(ms); print('tiles C=%d ms %.3f terms/s %.3e'%(C2,m,N*C2/m*1e3))
PY
