mkdir -p gpurun_out
python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -4
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from probayes_b200.engine import get_engine
from probayes_b200.cond_cov import CondCov
eng=get_engine(0); rng=np.random.default_rng(0)
d,Cg=64,65536
A=rng.standard_normal((d,d)); cov=A@A.T/d+np.eye(d); mean=rng.standard_normal(d)
cc=CondCov(mean,cov,np.tile([-10.,10.],(d,1)))
st=eng.to_device(np.tile(mean[:,None],(1,Cg)))
for sweeps,wp in [(4,True),(4,False),(16,False)]:
    ms=[]
    for _ in range(4):
        eng.gibbs_mvn(st,cc,sweeps*d,thin=d,seed=5,want_prob=wp); ms.append(eng.last_kernel_ms())
    print('gibbs d=64 C=65536 sweeps',sweeps,'want_prob',wp,'ms/sweep %.4f'%(np.median(ms[1:])/sweeps), 'coord updates/s %.3e'%(Cg*d*sweeps/(np.median(ms[1:])*1e-3)))
PY
