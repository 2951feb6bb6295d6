"""Serialisation of PDs: the reference's dict form and its HDF5 file
layout -- probayes/pd_utils.py:433-553, pd.py:698-703, distribution.py:286-290.

``serialise(*pds)`` / ``deserialise(dict)`` need nothing beyond numpy.  The four file
functions need ``h5py``, which this image does not have: they import it on first use and
raise ``ImportError`` otherwise (tests drive them through a dict-backed stand-in and compare
the stored structure with what the reference's own writer stores through the same stand-in).
A device-backed PD is copied to the host by ``PD.prob`` when it is serialised -- an explicit
export, not a compute path.

File layout (one HDF5 group per distribution, named by its short name, e.g. 'mu,sigma|x'):
  <key>      array values as they are; a set-valued key {n} as a zero-size byte array of
             shape (n, 0) (pd_utils.py:489-491)
  prob       the probability array; pscale: the scale (1.0 or 0j)
  attrs      a scalar dataset holding len(dims) whose HDF5 attributes are the dims (None
             stored as the string 'None') plus 'order', the key order
Auxiliary arrays go to groups whose names contain a space (pd_utils.py:466-477).
"""
import collections
import numpy as np

from .pd import PD


def _h5py():
    msg = ("probayes_b200.serial: the HDF5 functions need h5py, which is not installed here; "
           "serialise() / deserialise() work without it")
    try:
        import h5py
    except ImportError as e:
        raise ImportError(msg) from e
    if not hasattr(h5py, 'File'):                  # an empty import stub, not the library
        raise ImportError(msg)
    return h5py


def serialise(*args):
    """{short name: {key: value, ..., 'attrs': dims, 'prob': prob, 'pscale': pscale}}."""
    out = {}
    for arg in args:
        if not isinstance(arg, PD):
            raise TypeError("Unrecogised type to serialise: {}".format(type(arg)))
        out.update(arg.serialise())
    return out


def deserialise(serialised):
    assert isinstance(serialised, dict), \
        "Dict-type serialised input expected, not {}".format(type(serialised))
    dists = []
    for name, d in serialised.items():
        d = dict(d)
        dims = d.pop('attrs') if 'attrs' in d else {}
        prob = d.pop('prob') if 'prob' in d else None
        pscale = d.pop('pscale') if 'pscale' in d else None
        if pscale is not None:
            pscale = np.atleast_1d(pscale).tolist()[0]
        dists.append(PD(name, d, dims=dims, prob=prob, pscale=pscale))
    return tuple(dists)


def write_serialised(path, serialised, aux_dict=None):
    assert isinstance(serialised, dict), \
        "Dict-type serialised input expected, not {}".format(type(serialised))
    aux_dict = aux_dict or {}
    for key, val in aux_dict.items():
        assert isinstance(key, str) and ' ' in key, \
            "Aux dict must keyed by space-containing string, found {}".format(key)
        assert isinstance(val, dict), \
            "Aux dict must be a nested dictionary of dicts - found {}".format(type(val))
        for subkey, subval in val.items():
            assert isinstance(subkey, str), "Aux dict subkeys must be str, found {}".format(subkey)
            assert isinstance(subval, np.ndarray), \
                "Aux dict subvals must be NumPy arrays, found {}".format(type(subval))
    with _h5py().File(path, 'w', libver='latest') as f:
        for name, d in serialised.items():
            grp = f.create_group(name)
            for key, val in d.items():
                if key == 'attrs':
                    attrs = collections.OrderedDict((k, 'None' if v is None else v)
                                                    for k, v in val.items())
                    grp[key] = np.array(len(attrs))
                    grp[key].attrs.update(attrs)
                    grp[key].attrs.update({'order': list(attrs.keys())})
                elif isinstance(val, np.ndarray):
                    grp[key] = val
                elif isinstance(val, set):
                    grp[key] = np.zeros(sorted(val) + [0], dtype='S')
                else:
                    grp[key] = np.array(val)
        for name, arrays in aux_dict.items():
            grp = f.create_group(name)
            for key, val in arrays.items():
                grp[key] = val


def read_serialised(path):
    serialised, aux = {}, {}
    with _h5py().File(path, 'r', libver='latest') as f:
        for name in f.keys():
            grp = f[name]
            if ' ' in name:
                aux[name] = {key: np.array(val) for key, val in grp.items()}
                continue
            d, attrs, order = collections.OrderedDict(), None, None
            for key, val in grp.items():
                if key == 'attrs':
                    attrs = dict(val.attrs)
                    order = attrs.pop('order', None)
                    continue
                arr = np.array(val)
                d[key] = set(arr.shape[:-1]) if arr.dtype.kind == 'S' else arr
            if attrs is not None:
                order = [str(k) for k in (order if order is not None else attrs.keys())]
                dims = collections.OrderedDict()
                for k in order:
                    v = attrs[k]
                    dims[k] = None if (isinstance(v, (str, bytes)) and v in ('None', b'None')) \
                        else int(v)
                ordered = collections.OrderedDict((k, d[k]) for k in order)
                ordered.update((k, v) for k, v in d.items() if k not in ordered)
                d = ordered
                d['attrs'] = dims
            serialised[name] = d
    return serialised, aux


def write_dist(path, *args):
    return write_serialised(path, serialise(*args))


def read_dist(path):
    return deserialise(read_serialised(path)[0])
