"""Parity at BASELINE.json's FULL sizes (GPU, through the C ABI): the headline C2 walk
replayed step for step by the C oracle, the C4 grid at N = 10^5 on sampled cells, the C5
Gibbs configuration replayed on sampled chains.  Together they cost a few seconds."""
import numpy as np
import pytest
from conftest import relerr
from gpu_util import engine, dev, host
from oracle import np_oracle as o
from oracle import philox

pytestmark = pytest.mark.gpu
COV = np.array([[2.0, 1.2], [1.2, 2.0]])
NNI = -1.7976931348623157e+308


@pytest.mark.parametrize("variant", [0, 2])
def test_c2_full_size_replay(variant):
    """Config C2 exactly as bench.py runs it -- 4096 chains x 10^4 steps, native Philox RNG,
    log accept rule, every step recorded -- against the C restatement of the same walk
    (oracle/c, same Philox stream, glibc log / sincos instead of the device's table math):
    every chain's accept count identical, trajectories <= 1e-11, densities <= 1e-11.
    variant 0 = the default (whitened-decision) kernel, 2 = the exact-arithmetic one."""
    from oracle.c import liboracle as lo
    eng = engine()
    C, T, seed = 4096, 10000, 20261018
    init = np.tile(np.array([0., 1.]), (C, 1))
    lo.use_all_cores()
    ref = lo.mh_mvn_walk(init, [0., 0.], COV, T, seed, accept="log", record=True)
    state = dev(eng, init.T)
    out = eng.mh_mvn(state, [0., 0.], COV, T, seed=seed, accept="log", variant=variant)
    eng.sync()
    acc = host(out["accept_count"])
    assert np.array_equal(acc, ref["accept_count"]), \
        "%d chains differ in accept count" % int((acc != ref["accept_count"]).sum())
    x = host(out["x"])
    assert x.shape == ref["x"].shape == (T, 2, C)
    err = 0.0
    for t0 in range(0, T, 1000):                     # blockwise: no 1 GB temporaries
        err = max(err, float(np.abs(x[t0:t0 + 1000] - ref["x"][t0:t0 + 1000]).max()))
    assert err <= 1e-11, err
    assert relerr(host(out["prob"]), ref["prob"]) <= 1e-11
    assert np.abs(host(state) - ref["final"].T).max() <= 1e-11
    # moments of the whole run (Monte Carlo error ~ 1e-3 at 4e7 correlated draws)
    flat = x[1000:].transpose(1, 0, 2).reshape(2, -1)
    assert np.abs(flat.mean(axis=1)).max() < 5e-3
    assert np.abs(np.cov(flat) - COV).max() < 1e-2


def test_c4_full_size_sampled_cells():
    """Config C4: 4096 x 4096 grid over N = 10^5 observations.  64 sampled cells of the
    log-joint against the C oracle (<= 1e-12 relative), the posterior normalised to 1e-10,
    marginals = row / column sums of the posterior, sampled posterior cells against
    log-joint - log-sum-exp of the (device) log-joint computed on the host in long double."""
    from oracle.c import liboracle as lo
    eng = engine()
    rng = np.random.default_rng(7)
    N, M, S = 100_000, 4096, 4096
    data = rng.normal(50., 10., N)
    mu = o.uniform_grid(40, 60, M, True, True)
    sg = np.exp(o.uniform_grid(np.log(5), np.log(20), S, True, True))
    lpm, lps = np.full(M, -np.log(20.)), np.full(S, -np.log(np.log(4.)))
    lj = eng.grid_norm_logjoint(dev(eng, data), dev(eng, mu), dev(eng, sg), dev(eng, lpm),
                                dev(eng, lps))
    eng.sync()
    ljh = host(lj)
    rows = np.concatenate([[0, M - 1], rng.choice(M, 6, replace=False)])
    cols = np.concatenate([[0, S - 1], rng.choice(S, 6, replace=False)])
    # include the neighbourhood of the mode, where the posterior mass is
    im, js = np.unravel_index(np.argmax(ljh), ljh.shape)
    rows[2:4], cols[2:4] = [im, min(im + 1, M - 1)], [js, min(js + 1, S - 1)]
    lo.use_all_cores()
    want = lo.grid_norm_logjoint(data, mu[rows], sg[cols], lpm[rows], lps[cols])
    assert relerr(ljh[np.ix_(rows, cols)], want) <= 1e-12
    r = eng.grid_conditionalise(lj)
    eng.sync()
    post = host(r["post"])
    lin = np.where(post == NNI, 0.0, np.exp(np.maximum(post, -745.)))
    assert abs(lin.sum() - 1.0) <= 1e-10
    # normaliser on the host in extended precision from the device log-joint
    gmax = ljh.max()
    lse = gmax + float(np.log(np.exp((ljh - gmax).astype(np.longdouble)).sum()))
    sel = post[np.ix_(rows, cols)]
    ok = sel != NNI
    assert np.abs(sel[ok] - (ljh[np.ix_(rows, cols)][ok] - lse)).max() <= 1e-12 * abs(gmax)
    mm, ms = host(r["marg_mu"]), host(r["marg_sigma"])
    ok = mm > -700
    assert relerr(np.exp(mm[ok]), lin.sum(axis=1)[ok]) <= 1e-10
    ok = ms > -700
    assert relerr(np.exp(ms[ok]), lin.sum(axis=0)[ok]) <= 1e-10


def test_c5_full_size_sampled_chain_replay():
    """Config C5: d = 64 Gibbs, 65536 chains, native RNG.  128 sampled chains are replayed
    on the CPU from the same Philox uniforms through the numpy restatement of CondCov
    (scipy ndtri); moments of all chains within Monte Carlo error."""
    from probayes_b200.cond_cov import CondCov
    eng = engine()
    d, C, sweeps, seed = 64, 65536, 3, 97531
    rng = np.random.default_rng(0)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    lims = np.tile([-10., 10.], (d, 1))
    cc = CondCov(mean, cov, lims)
    T = sweeps * d
    state = dev(eng, np.tile(mean[:, None], (1, C)))
    out = eng.gibbs_mvn(state, cc, T, thin=d, seed=seed, want_prob=True, log_pscale=True)
    eng.sync()
    X = host(out["x"])                                        # [sweeps, d, C]
    sel = np.sort(rng.choice(C, 128, replace=False))
    t = np.arange(T, dtype=np.uint64)[:, None]
    R = philox.gibbs_uniforms(seed, t, sel.astype(np.uint64)[None, :])
    ref = o.gibbs_mvn_walk(np.tile(mean, (128, 1)), R, mean, cov, lims, log_pscale=True)
    want = np.transpose(ref["x"], (0, 2, 1))[d - 1::d]        # [sweeps, d, 128]
    err = float(np.abs(X[:, :, sel] - want).max())
    assert err <= 1e-12, err
    assert relerr(host(out["prob"])[:, sel], ref["prob"][d - 1::d]) <= 1e-12
    # every chain: finite, inside the box; the first coordinate's pooled mean moves from
    # mean[0] by less than a few standard errors of an iid sample (the sampler started AT
    # the mean, so after 3 sweeps it is still within sd / sqrt(C) * O(1))
    assert np.isfinite(X).all() and np.abs(X).max() < 10.0
    se = np.sqrt(np.diag(cov) / C)
    assert (np.abs(X[-1].mean(axis=1) - mean) < 6 * se).all()
