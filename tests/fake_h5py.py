"""A dict-backed stand-in for the few h5py calls the reference's HDF5 serialisers make
(pd_utils.py:464-545): File as a context manager, create_group, dataset assignment / read
back through np.array, .attrs, .keys() / .items().  h5py is absent from this image; the
stand-in lets the reference's writer and ours be driven side by side and their output
compared structurally.  Files live in the module-level STORE, keyed by path."""
import collections
import numpy as np

STORE = {}


class Dataset:
    def __init__(self, value):
        self.value = np.array(value)
        self.attrs = {}

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


class Group(collections.OrderedDict):
    def create_group(self, name):
        g = Group()
        collections.OrderedDict.__setitem__(self, name, g)
        return g

    def __setitem__(self, key, value):
        collections.OrderedDict.__setitem__(self, key, value if isinstance(value, (Group, Dataset))
                                            else Dataset(value))


class File(Group):
    def __init__(self, path, mode='r', libver=None):
        super().__init__()
        self._path, self._mode = path, mode
        if mode == 'r':
            if path not in STORE:
                raise OSError("no such file: " + str(path))
            self.update(STORE[path])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self._mode != 'r':
            STORE[self._path] = Group(self)
        return False


def dump(path):
    """{group: {key: (dtype str, shape, values list, attrs dict)}} of a stored file."""
    out = collections.OrderedDict()
    for gname, grp in STORE[path].items():
        out[gname] = collections.OrderedDict()
        for key, ds in grp.items():
            attrs = {k: (list(v) if isinstance(v, (list, tuple, np.ndarray)) else v)
                     for k, v in ds.attrs.items()}
            out[gname][key] = (ds.value.dtype.kind, tuple(ds.value.shape),
                               np.asarray(ds.value).tolist() if ds.value.dtype.kind != 'S' else None,
                               attrs)
    return out
