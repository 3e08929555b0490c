"""
Container files of the diagnostics (fbpic_b200/diags.py).

The reference writes openPMD files through h5py (fbpic/openpmd_diag/generic_diag.py:14,117).  `open_file` returns
  * an `h5py.File` when h5py can be imported: the output then is a real openPMD/HDF5 series, `hdf5/data%08d.h5`;
  * otherwise (this build image has no HDF5 library) a `File` of this module: the same tree of groups, datasets and
    attributes behind the part of the h5py API the diagnostics use, stored as a NumPy archive `hdf5/data%08d.npz`
    whose keys are the HDF5 paths: '<dataset path>' -> array, '<node path>@<attribute>' -> value and
    '<group path>/' -> empty array for a group without children (openPMD constant records).

`read_tree(path)` gives that flat dictionary for either kind of file.
"""
import os
import numpy as np


def have_h5py():
    try:
        import h5py          # noqa: F401
        return True
    except ImportError:
        return False


def open_file(path_without_extension, mode='a'):
    if have_h5py():
        import h5py
        return h5py.File(path_without_extension + '.h5', mode)
    return File(path_without_extension + '.npz', mode)


def existing_file(path_without_extension):
    for ext in ('.h5', '.npz'):
        if os.path.exists(path_without_extension + ext):
            return path_without_extension + ext
    return None


class _Node(object):
    def __init__(self, name, root):
        self.name, self._root, self.attrs = name, (self if root is None else root), {}


class Dataset(_Node):
    def __init__(self, name, root, array):
        _Node.__init__(self, name, root)
        self._a = array

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self._a[idx] = value

    def __len__(self):
        return len(self._a)

    def resize(self, size, axis=0):
        shape = list(self._a.shape)
        shape[axis] = size
        new = np.zeros(shape, dtype=self._a.dtype)
        keep = tuple(slice(0, min(a, b)) for a, b in zip(shape, self._a.shape))
        new[keep] = self._a[keep]
        self._a = new


class Group(_Node):
    def __init__(self, name, root):
        _Node.__init__(self, name, root)
        self._children = {}

    def _descend(self, path, create):
        node = self._root if path.startswith('/') else self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group) or part not in node._children:
                if not create:
                    raise KeyError("'%s' is not in %s" % (path, self.name))
                node._children[part] = Group(node.name.rstrip('/') + '/' + part, self._root)
            node = node._children[part]
        return node

    def _split(self, path):
        head, _, leaf = path.rstrip('/').rpartition('/')
        if path.startswith('/') and not head:
            return self._root, leaf
        return (self._descend(head, True) if head else self), leaf

    def require_group(self, path):
        return self._descend(path, True)

    create_group = require_group

    def create_dataset(self, path, shape=None, dtype=None, data=None, **kw):
        parent, leaf = self._split(path)
        if leaf in parent._children:
            raise ValueError('Unable to create dataset (name already exists): %s' % path)
        array = np.zeros(shape, dtype=dtype) if data is None else np.array(data, dtype=dtype)
        d = Dataset(parent.name.rstrip('/') + '/' + leaf, self._root, array)
        parent._children[leaf] = d
        return d

    def require_dataset(self, path, shape, dtype, **kw):
        if path in self:
            d = self[path]
            if tuple(d.shape) != tuple(shape):
                raise TypeError('Shapes do not match (existing %s vs new %s)' % (d.shape, tuple(shape)))
            return d
        return self.create_dataset(path, shape, dtype)

    def __getitem__(self, path):
        return self._descend(path, False)

    def __contains__(self, path):
        try:
            self._descend(path, False)
            return True
        except KeyError:
            return False

    def __delitem__(self, path):
        parent, leaf = self._split(path)
        del parent._children[leaf]

    def keys(self):
        return self._children.keys()


class File(Group):
    """mode 'a': read/write, created if missing; 'w': truncate; 'r': read only."""

    def __init__(self, filename, mode='a'):
        Group.__init__(self, '/', None)
        self.filename, self.mode = filename, mode
        if mode in ('a', 'r', 'r+') and os.path.exists(filename):
            with np.load(filename, allow_pickle=False) as z:
                for key in (k for k in z.files if '@' not in k):
                    if key.endswith('/'):
                        self._descend(key, True)
                    else:
                        parent, leaf = self._split(key)
                        parent._children[leaf] = Dataset(key, self, z[key])
                for key in (k for k in z.files if '@' in k):
                    path, attr = key.rsplit('@', 1)
                    value = z[key]
                    self._descend(path, True).attrs[attr] = value[()] if value.ndim == 0 else value
        elif mode in ('r', 'r+'):
            raise OSError("Unable to open file (no such file: '%s')" % filename)

    def flat(self):
        out = {}

        def visit(node):
            for k, v in node.attrs.items():
                out['%s@%s' % (node.name, k)] = np.bytes_(v) if isinstance(v, (bytes, str)) and not isinstance(v, np.generic) \
                    else v
            if isinstance(node, Dataset):
                out[node.name] = node._a
            else:
                if not node._children and node is not self:
                    out[node.name + '/'] = np.zeros(0)
                for child in node._children.values():
                    visit(child)
        visit(self)
        return out

    def close(self):
        if self.mode != 'r':
            tmp = self.filename + '.tmp.npz'
            np.savez(tmp, **self.flat())
            os.replace(tmp, self.filename)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_tree(filename):
    """{'<dataset path>': array, '<node path>@<attribute>': value, '<empty group path>/': empty array}"""
    if filename.endswith('.npz'):
        with np.load(filename, allow_pickle=False) as z:
            return {k: (z[k][()] if z[k].ndim == 0 else z[k]) for k in z.files}
    import h5py
    out = {}

    def visit(name, node):
        path = '/' + name.strip('/') if name else '/'
        for k, v in node.attrs.items():
            out['%s@%s' % (path, k)] = v
        if hasattr(node, 'shape'):
            out[path] = node[...]
        elif len(node.keys()) == 0 and path != '/':
            out[path + '/'] = np.zeros(0)
    with h5py.File(filename, 'r') as f:
        visit('', f)
        f.visititems(visit)
    return out


def npz_to_hdf5(npz_path, h5_path=None):
    """Rewrite one `.npz` archive of this module as the HDF5 file it stands for (needs h5py): same groups, datasets
    and attributes, i.e. a regular openPMD file.  `python -m fbpic_b200.openpmd_store <dir or file> ...` converts whole
    diagnostics directories after a run on a machine without h5py."""
    import h5py
    h5_path = h5_path or npz_path[:-4] + '.h5'
    tree = read_tree(npz_path)
    with h5py.File(h5_path, 'w') as f:
        for key in sorted(k for k in tree if '@' not in k):
            if key.endswith('/'):
                f.require_group(key)
            else:
                f.create_dataset(key, data=tree[key])
        for key in (k for k in tree if '@' in k):
            path, attr = key.rsplit('@', 1)
            (f if path == '/' else f[path]).attrs[attr] = tree[key]
    return h5_path


if __name__ == '__main__':
    import sys
    for target in sys.argv[1:]:
        files = [os.path.join(root, n) for root, _, names in os.walk(target) for n in names if n.endswith('.npz')] \
            if os.path.isdir(target) else [target]
        for path in sorted(files):
            print(npz_to_hdf5(path))
