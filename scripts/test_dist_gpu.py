"""Multi-GPU check of the sharded paths with real NCCL (run under torchrun):
  * MH chains sharded over ranks == the same global chains on one GPU (Philox keyed
    on the global chain id), R-hat from all-reduced summaries;
  * DGEI mu-row slabs: normaliser / marginals agree with the single-GPU result;
  * OMC random samples sharded over ranks: same global Philox stream, posterior means
    from all-reduced sums agree with the single-GPU result.
Prints PASS lines on rank 0; exits non-zero on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from probayes_b200.engine import get_engine
from probayes_b200 import dist as pd_

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
eng = get_engine(local)
COV = [[2., 1.2], [1.2, 2.]]
C, T, seed = 4096, 500, 77
init = np.random.default_rng(1).standard_normal((2, C))
start, count = pd_.shard_range(C, rank, world)
out = eng.mh_mvn(eng.to_device(init[:, start:start + count]), [0., 0.], COV, T, seed=seed,
                 chain0=start, accept="log")
st = pd_.allreduce_chain_stats(eng.chain_stats(out["stat_sum"], out["stat_sumsq"], T))
rh = pd_.rhat_from_stats(st, T)
xs = [torch.empty((T, 2, pd_.shard_range(C, r, world)[1]), dtype=torch.float64, device="cuda")
      for r in range(world)]
dist.all_gather(xs, out["x"].contiguous())
ok = True
if rank == 0:
    full = eng.mh_mvn(eng.to_device(init), [0., 0.], COV, T, seed=seed, accept="log")
    same = torch.equal(torch.cat(xs, dim=2), full["x"])
    st1 = eng.chain_stats(full["stat_sum"], full["stat_sumsq"], T)
    rh1 = pd_.rhat_from_stats(st1, T)
    print("chains sharded == single GPU:", same, "rhat", rh, rh1)
    ok &= same and np.allclose(rh, rh1, rtol=1e-12)
# DGEI
rng = np.random.default_rng(3)
N, M, S = 5000, 301, 257
data = rng.normal(50., 10., N)
mu = np.linspace(40, 60, M + 2)[1:-1]
sg = np.exp(np.linspace(np.log(5), np.log(20), S + 2)[1:-1])
lpm, lps = np.full(M, -np.log(20.)), np.full(S, -np.log(np.log(4.)))
r = pd_.dgei_sharded(eng, data, mu, sg, lpm, lps)
posts = [torch.empty((pd_.shard_range(M, q, world)[1], S), dtype=torch.float64, device="cuda")
         for q in range(world)]
dist.all_gather(posts, r["post"].contiguous())
if rank == 0:
    lj = eng.grid_norm_logjoint(eng.to_device(data), eng.to_device(mu), eng.to_device(sg),
                                eng.to_device(lpm), eng.to_device(lps))
    w = eng.grid_conditionalise(lj)
    post = torch.cat(posts)
    keep = w["post"] > -1e300
    e1 = float((post[keep] - w["post"][keep]).abs().max())
    e2 = float((r["marg_mu"] - w["marg_mu"]).abs().max())
    e3 = float((r["marg_sigma"] - w["marg_sigma"]).abs().max())
    same_mask = bool(torch.equal(post > -1e300, keep))
    print("dgei sharded vs single: post %.2e marg_mu %.2e marg_sigma %.2e mask %s" % (e1, e2, e3, same_mask))
    ok &= same_mask and e1 < 1e-9 and e2 < 1e-9 and e3 < 1e-9
# OMC random sampling, samples sharded over ranks
lims = np.array([[40., 60.], [5., 20.]]); logu = np.array([0, 1]); ex = np.ones((2, 2), int)
Tn = 100003
ro = pd_.omc_sharded(eng, data[:300], lims, ex, logu, Tn, seed=21)
ths = [torch.empty((2, pd_.shard_range(Tn, q, world)[1]), dtype=torch.float64, device="cuda")
       for q in range(world)]
dist.all_gather(ths, ro["theta"].contiguous())
if rank == 0:
    th1 = eng.box_sample(lims, logu, Tn, seed=21)
    same = torch.equal(torch.cat(ths, dim=1), th1)
    lp1 = eng.normreg_logjoint(th1, eng.to_device(data[:300]), None, lims, ex, logu)
    w1 = eng.grid_conditionalise(lp1.reshape(1, -1))
    s1 = eng.expectation_sums(w1["post"].reshape(-1), True, None, th1).cpu().numpy()
    e1 = s1[1:] / s1[0]
    print("omc sharded == single GPU draws:", same, "expectation", ro["expectation"], e1)
    ok &= same and np.allclose(ro["expectation"], e1, rtol=1e-11)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("PASS" if ok else "FAIL")
sys.exit(0 if int(flag.item()) else 1)
