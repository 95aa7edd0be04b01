"""Stand-in for h5py: full_model.py reads pretrained weights with `h5py.File(path, 'r')[key][:]`.  Files are
in-memory dicts registered under a path name (REGISTRY[path] = {key: array})."""
REGISTRY = {}


class File(object):

  def __init__(self, path, mode='r'):
    self._d = REGISTRY[path]

  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False

  def __getitem__(self, key):
    return self._d[key]

  def __contains__(self, key):
    return key in self._d

  def close(self):
    pass
