"""Stand-in for `h5py` (not installed in the build container; no HDF5 library in the image).  TEST INFRASTRUCTURE.

It implements the small part of the h5py API that the reference's openPMD writers use
(fbpic/openpmd_diag/{generic_diag,field_diag,particle_diag,boosted_field_diag,boosted_particle_diag}.py:
`File(path, mode)`, `require_group`, `require_dataset`, `create_dataset`, item access by path, `in`, `del`,
`.attrs`, dataset slicing and `resize`, `close`), so that the UNMODIFIED reference diagnostics run here and the tree
they write (groups, datasets, attributes) can be harvested by `oracle/gen_golden_ext.py` as a fixture.  A "file" is a
pickled tree at the path the reference chose -- it is not an HDF5 file.  The laser package of the reference also
imports h5py at module level (fbpic/lpa_utils/laser/laser_profiles.py:10); nothing more than the import is needed
there."""
import os
import pickle
import numpy as np

__version__ = '0.0-shim'


class _Node(object):
    def __init__(self, name):
        self.name = name
        self.attrs = {}


class Dataset(_Node):
    def __init__(self, name, shape, dtype, maxshape=None):
        _Node.__init__(self, name)
        self._a = np.zeros(shape, dtype=dtype)
        self.maxshape = maxshape

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self._a[idx] = value

    def __len__(self):
        return len(self._a)

    def resize(self, size, axis=None):
        shape = list(self._a.shape)
        if axis is None:
            shape = list(size)
        else:
            shape[axis] = size
        new = np.zeros(shape, dtype=self._a.dtype)
        sl = tuple(slice(0, min(o, n)) for o, n in zip(self._a.shape, shape))
        new[sl] = self._a[sl]
        self._a = new


class Group(_Node):
    def __init__(self, name):
        _Node.__init__(self, name)
        self._children = {}

    def _walk(self, path, create):
        node = self
        if path.startswith('/'):
            node = self._root
        parts = [p for p in path.split('/') if p]
        for i, p in enumerate(parts):
            if p not in node._children:
                if not create:
                    raise KeyError(path)
                g = Group(node.name.rstrip('/') + '/' + p)
                g._root = node._root
                node._children[p] = g
            node = node._children[p]
        return node

    def _parent_and_leaf(self, path):
        parts = [p for p in path.split('/') if p]
        base = self._root if path.startswith('/') else self
        return base._walk('/'.join(parts[:-1]), True), parts[-1]

    def require_group(self, path):
        return self._walk(path, True)

    create_group = require_group

    def create_dataset(self, path, shape=None, dtype=None, data=None, maxshape=None, **kw):
        parent, leaf = self._parent_and_leaf(path)
        if leaf in parent._children:
            raise ValueError('name already exists: %s' % path)
        if data is not None:
            data = np.asarray(data, dtype=dtype)
            shape = data.shape
            dtype = data.dtype
        d = Dataset(parent.name.rstrip('/') + '/' + leaf, shape, dtype, maxshape)
        if data is not None:
            d._a[...] = data
        parent._children[leaf] = d
        return d

    def require_dataset(self, path, shape, dtype, **kw):
        if path in self:
            d = self[path]
            assert tuple(d.shape) == tuple(shape)
            return d
        return self.create_dataset(path, shape, dtype, **kw)

    def __getitem__(self, path):
        return self._walk(path, False)

    def __contains__(self, path):
        try:
            self._walk(path, False)
            return True
        except KeyError:
            return False

    def __delitem__(self, path):
        parent, leaf = self._parent_and_leaf(path)
        del parent._children[leaf]

    def keys(self):
        return self._children.keys()

    def items(self):
        return self._children.items()

    def visititems(self, func):
        for k, child in self._children.items():
            func(child.name.lstrip('/'), child)
            if isinstance(child, Group):
                child.visititems(func)


class File(Group):
    def __init__(self, path, mode='r', **kw):
        Group.__init__(self, '/')
        self._root = self
        self.filename = path
        self.mode = mode
        if mode in ('r', 'r+', 'a') and os.path.exists(path):
            with open(path, 'rb') as f:
                saved = pickle.load(f)
            self.attrs, self._children = saved.attrs, saved._children
            _reroot(self, self)
        elif mode in ('r', 'r+'):
            raise OSError('Unable to open file %s' % path)

    def close(self):
        if self.mode != 'r':
            with open(self.filename, 'wb') as f:
                pickle.dump(self, f)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _reroot(node, root):
    node._root = root
    for c in getattr(node, '_children', {}).values():
        if isinstance(c, Group):
            _reroot(c, root)


def flatten(path):
    """{'<dataset path>': array, '<node path>@<attribute>': value} of a shim file (for the fixtures)."""
    out = {}

    def visit(node):
        for k, v in node.attrs.items():
            out['%s@%s' % (node.name, k)] = v
        if isinstance(node, Dataset):
            out[node.name] = node._a
        else:
            if not node._children:
                out[node.name.rstrip('/') + '/'] = np.zeros(0)
            for c in node._children.values():
                visit(c)
    visit(File(path, 'r'))
    return out
