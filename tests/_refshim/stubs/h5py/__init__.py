"""
TESTS ONLY -- an in-memory stand-in for the small part of ``h5py`` the reference's
model / particle / ICs code uses (``File``, groups, datasets, ``attrs``), so that
``ClusterModel.write_model_to_h5 / from_h5_file``, ``ClusterParticles.write_particles /
from_file`` and ``ClusterICs`` run unmodified.  "Files" live in a process-wide
dict keyed by path; nothing touches the disk.
"""
import numpy as np

_STORE = {}


def clear():
    _STORE.clear()


class Dataset:
    def __init__(self, data):
        self._data = np.array(data)
        self.attrs = {}

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def size(self):
        return self._data.size

    def __getitem__(self, item):
        if item == ():
            return self._data[()]
        return self._data[item]

    def __setitem__(self, item, value):
        self._data[item] = value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __len__(self):
        return len(self._data)


class Group:
    def __init__(self):
        self._items = {}
        self.attrs = {}

    def _walk(self, key, create=False):
        node = self
        parts = [p for p in key.split("/") if p]
        for p in parts[:-1]:
            if p not in node._items:
                if not create:
                    raise KeyError(key)
                node._items[p] = Group()
            node = node._items[p]
        return node, parts[-1]

    def create_group(self, name):
        node, leaf = self._walk(name, create=True)
        if leaf in node._items:
            raise ValueError(f"group {name} exists")
        node._items[leaf] = Group()
        return node._items[leaf]

    def require_group(self, name):
        return self[name] if name in self else self.create_group(name)

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kw):
        node, leaf = self._walk(name, create=True)
        if leaf in node._items:
            raise ValueError(f"dataset {name} exists")
        if data is None:
            data = np.zeros(shape, dtype=dtype or "float64")
        elif dtype is not None:
            data = np.asarray(data).astype(dtype)
        node._items[leaf] = Dataset(data)
        return node._items[leaf]

    def __getitem__(self, key):
        node, leaf = self._walk(key)
        return node._items[leaf]

    def __delitem__(self, key):
        node, leaf = self._walk(key)
        del node._items[leaf]

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    def __iter__(self):
        return iter(self._items)

    def __len__(self):
        return len(self._items)


class File(Group):
    def __init__(self, name, mode="r", **kw):
        name = str(name)
        if mode in ("w", "w-", "x"):
            _STORE[name] = Group()
        elif mode in ("a", "r+"):
            if name not in _STORE:
                if mode == "r+":
                    raise OSError(f"no such in-memory file {name}")
                _STORE[name] = Group()
        elif name not in _STORE:
            raise OSError(f"no such in-memory file {name}")
        root = _STORE[name]
        self._items = root._items
        self.attrs = root.attrs
        self.filename = name

    def flush(self):
        pass

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
