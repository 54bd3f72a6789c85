"""h5py stand-in (read side) over the repo's pure-Python HDF5 subset reader, exposing what the
reference's weight loader touches: ``File`` as a context manager, ``attrs`` (``in`` / ``[]``, byte
strings like h5py returns for fixed-length string attributes), ``in`` / ``[]`` on groups, and
datasets convertible with ``np.asarray``."""
import numpy as np

from uplift_upsample_3dhpe_b200 import h5lite as _h5

__version__ = "h5lite-shim"


class _Attrs:
    def __init__(self, obj):
        self._a = obj.attrs

    def __contains__(self, k):
        return k in self._a

    def __getitem__(self, k):
        v = self._a[k]
        if isinstance(v, str):
            return v.encode("utf8")
        if isinstance(v, (list, tuple)) or (isinstance(v, np.ndarray) and v.dtype.kind in "OUS"):
            return np.array([s.encode("utf8") if isinstance(s, str) else s for s in v], dtype=object)
        return v

    def keys(self):
        return self._a.keys()


class _Dataset:
    def __init__(self, obj):
        self._o = obj

    def __array__(self, dtype=None, copy=None):
        a = self._o.read()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return self._o.read().shape


class _Group:
    def __init__(self, obj):
        self._o = obj
        self.attrs = _Attrs(obj)

    def __contains__(self, name):
        return name in self._o

    def __getitem__(self, path):
        o = self._o[path]
        return _Group(o) if o.is_group else _Dataset(o)


class File(_Group):
    def __init__(self, path, mode="r"):
        assert mode == "r", "h5py shim is read-only"
        super().__init__(_h5.H5File(path))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
