"""the dict form of PDs and the HDF5 file layout
(probayes/pd_utils.py:433-553) against the live reference's own output
(tests/golden/pd_serialise.npz, written by oracle/gen_golden.py).  h5py is absent from the
image, so both the reference's writer (when the fixture was made) and ours (here) store
through tests/fake_h5py.py and the stored structures are compared."""
import collections
import json
import os
import sys
import numpy as np
import pytest
from conftest import load_golden

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_h5py  # noqa: E402
import probayes_b200 as pb  # noqa: E402


def _pds(g):
    N = len(g["data"])
    mu, sigma = g["mu"], g["sigma"]
    od = collections.OrderedDict
    joint = pb.PD('mu,sigma,x', od([('mu', mu[:, None]), ('sigma', sigma[None, :]), ('x', {N})]),
                  dims=od([('mu', 0), ('sigma', 1), ('x', None)]), prob=g["joint_prob"],
                  pscale='log')
    post = pb.PD('mu,sigma|x', od([('mu', mu[:, None]), ('sigma', sigma[None, :]), ('x', {N})]),
                 dims=od([('mu', 0), ('sigma', 1), ('x', None)]), prob=g["posterior_prob"],
                 pscale='log')
    pmu = pb.PD('mu|x', od([('mu', mu), ('x', {N})]), dims=od([('mu', 0), ('x', None)]),
                prob=g["post_mu_prob"], pscale=1.)
    return dict(joint=joint, posterior=post, post_mu=pmu)


def test_dict_form_matches_the_reference():
    g = load_golden("pd_serialise")
    for tag, pd in _pds(g).items():
        ser = pb.serialise(pd)
        (name, d), = ser.items()
        assert name == str(g[tag + "_name"])
        assert list(d.keys()) == json.loads(str(g[tag + "_keys"]))
        assert {k: v for k, v in d['attrs'].items()} == json.loads(str(g[tag + "_dims"]))
        assert np.iscomplexobj(d['pscale']) == bool(g[tag + "_pscale_is_log"])
        assert np.array_equal(d['prob'], g[tag + "_prob"])
        back, = pb.deserialise(ser)
        assert back.name == str(g[tag + "_back_name"])
        assert np.array_equal(back.prob, pd.prob) and back.pscale == pd.pscale
        assert list(back.dims.items()) == list(pd.dims.items())
    with pytest.raises(TypeError):
        pb.serialise({'not': 'a PD'})


def test_file_layout_matches_the_reference(monkeypatch):
    monkeypatch.setitem(sys.modules, 'h5py', fake_h5py)
    g = load_golden("pd_serialise")
    for tag, pd in _pds(g).items():
        path = "ours_" + tag
        aux = {"aux data": {"obs": g["data"]}} if tag == "joint" else {}
        before = list(pd.dims.items())
        pb.write_serialised(path, pb.serialise(pd), aux)
        assert list(pd.dims.items()) == before            # the PD's dims are left alone
        ours = json.loads(json.dumps(fake_h5py.dump(path),
                                     default=lambda o: o.item() if hasattr(o, "item") else str(o)))
        assert ours == json.loads(str(g[tag + "_file"]))
        ser, auxr = pb.read_serialised(path)
        if aux:
            assert np.array_equal(auxr["aux data"]["obs"], g["data"])
        back, = pb.read_dist(path)
        assert back.name == str(g[tag + "_back_name"])
        assert np.array_equal(back.prob, g[tag + "_back_prob"])
        assert back.pscale == pd.pscale
        for k in ('mu', 'sigma'):
            if k in pd:
                assert np.array_equal(back[k], pd[k])
        assert back['x'] == {len(g["data"])}
    # write_dist / read_dist with several distributions in one file
    pds = _pds(g)
    pb.write_dist("ours_all", pds["joint"], pds["post_mu"])
    a, b = pb.read_dist("ours_all")
    assert np.array_equal(a.prob, pds["joint"].prob) and np.array_equal(b.prob, pds["post_mu"].prob)
    with pytest.raises(AssertionError):
        pb.write_serialised("bad", pb.serialise(pds["joint"]), {"nospace": {"a": np.zeros(2)}})


def test_file_functions_need_h5py():
    try:
        import h5py
        if hasattr(h5py, 'File'):
            pytest.skip("h5py is installed here")
    except ImportError:
        pass
    g = load_golden("pd_serialise")
    with pytest.raises(ImportError, match="h5py"):
        pb.write_dist("/tmp/never.h5", _pds(g)["post_mu"])
