"""The C restatement (oracle/c) agrees with the numpy restatement and the
golden fixtures; Philox matches the Random123 known-answer vectors."""
import numpy as np
import pytest
from conftest import load_golden, relerr
from oracle import np_oracle as o
from oracle import philox


@pytest.fixture(scope="module")
def lo():
    from oracle.c import build, liboracle
    build.build()
    return liboracle


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2,
            (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_u01_open_interval():
    z, f = np.uint64(0), np.uint64(0xffffffff)
    for lo_, hi_ in [(philox.u52(z, z), philox.u52(f, f)), (philox.u32(z), philox.u32(f)),
                     (philox.t44(z, z), philox.t44(f, f))]:
        assert 0.0 < lo_ < 1e-9 and 1.0 - 1e-9 < hi_ < 1.0


def test_c_mh_mvn_matches_numpy(lo):
    C, T, seed = 48, 150, 99
    cov = np.array([[2., 1.2], [1.2, 2.]])
    init = np.tile([0., 1.], (C, 1))
    for accept, logp in [("reference", False), ("log", True)]:
        r = lo.mh_mvn_walk(init, [0, 0], cov, T, seed, log_pscale=logp, accept=accept)
        Z = philox.normals(seed, T, C, 2)
        U = philox.thresholds(seed, T, C)
        ref = o.mh_mvn_walk(init, Z, U, [0, 0], cov, log_pscale=logp, accept=accept)
        assert np.array_equal(r["u"], ref["u"])
        assert np.abs(r["x"] - np.transpose(ref["x"], (0, 2, 1))).max() <= 1e-12
        assert relerr(r["prob"], ref["prob"]) <= 1e-12


def test_c_normreg_grid_gibbs_match_golden(lo):
    g = load_golden("mh_linreg")
    lj = lo.normreg_logjoint(g["x"], g["x_obs"], g["y_obs"], g["lims"], g["ex"], g["log_ufun"])
    assert relerr(lj, g["prob"]) <= 1e-12
    g = load_golden("mh_norm1d_metropolis")
    lj = lo.normreg_logjoint(g["x"], None, g["x_obs"], g["lims"], g["ex"], g["log_ufun"])
    assert relerr(lj, g["prob"]) <= 1e-12
    g = load_golden("dgei_small")
    M, S = len(g["mu"]), len(g["sigma"])
    lj = lo.grid_norm_logjoint(g["data"], g["mu"], g["sigma"], np.full(M, -np.log(20.)),
                               np.full(S, -np.log(np.log(4.))))
    assert relerr(lj, g["joint"]) <= 1e-12
    post, mm, ms = lo.grid_posterior(lj)
    assert relerr(post, g["posterior"]) <= 1e-12
    assert relerr(mm, g["marg_mu"]) <= 1e-12 and relerr(ms, g["marg_sigma"]) <= 1e-12
    g = load_golden("condcov_d64")
    x = lo.gibbs_mvn_walk(g["init"][None], g["mean"], g["coef"], g["stdv"], g["cdfs"],
                          len(g["runif"]), runif=g["runif"][:, None])
    assert np.abs(x[0] - g["x"][-1]).max() <= 1e-11


def test_gibbs_native_stream_c_vs_numpy(lo):
    """The C restatement's native Gibbs stream (two steps per Philox block: g and g + 4) is
    the one oracle/philox.py replays: both oracles give the same trajectory."""
    from probayes_b200.cond_cov import CondCov
    d, Cn, T, seed, step0, chain0 = 8, 5, 43, 77, 16, 3
    rng = np.random.default_rng(4)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    lims = np.tile([-10., 10.], (d, 1))
    cc = CondCov(mean, cov, lims)
    init = np.tile(mean, (Cn, 1))
    x = lo.gibbs_mvn_walk(init, mean, cc.coef_matrix(), cc.stdv, cc.cdfs, T, seed=seed, step0=step0,
                          chain0=chain0)
    t = (np.arange(T, dtype=np.uint64) + np.uint64(step0))[:, None]
    c = (np.arange(Cn, dtype=np.uint64) + np.uint64(chain0))[None, :]
    R = philox.gibbs_uniforms(seed, t, c)
    assert R.shape == (T, Cn) and (R > 0).all() and (R < 1).all()
    # steps g and g + 4 come from one block, different words
    assert not np.array_equal(R[0], R[4])
    ref = o.gibbs_mvn_walk(init, R, mean, cov, lims, start=step0)
    assert np.abs(x - ref["x"][-1]).max() <= 1e-11
