"""structs.numpy stand-in: Table (struct of arrays sharing a leading shape) and the
ordered tree traversals the reference's parameters.py relies on."""
import numpy as np
from .struct import Struct, struct


def _common_prefix(shapes):
  prefix = []
  for dims in zip(*shapes):
    if all(d == dims[0] for d in dims): prefix.append(dims[0])
    else: break
  return tuple(prefix)


class Table(Struct):
  def __init__(self, entries=None, **kw):
    super().__init__(entries, **kw)
    assert all(isinstance(v, np.ndarray) for v in self.values()), "Table: arrays only"
    self.__dict__  # noqa

  @staticmethod
  def create(**d): return Table(d)
  @staticmethod
  def build(d): return Table(d)
  @staticmethod
  def stack(tables, axis=0):
    t = tables[0]
    return Table({k: np.stack([x[k] for x in tables], axis=axis) for k in t.keys()})

  @property
  def _prefix(self): return _common_prefix([v.shape for v in self.values()])
  @property
  def _shape(self): return self._prefix
  @property
  def _size(self): return self._prefix[0]
  def _index(self, idx): return _index(self, idx)
  def _index_select(self, idx, axis=0):
    return Table({k: np.take(v, idx, axis=axis) for k, v in self.items()})
  def _sequence(self, axis=0):
    n = self._prefix[axis]
    return [self._index_select(i, axis=axis) for i in range(n)]
  def _sum(self, axis=0): return self._map(lambda a: a.sum(axis=axis))


def _index(t, idx):
  out = {k: v[idx] for k, v in t.items()}
  if all(isinstance(v, np.ndarray) for v in out.values()): return Table(out)
  return Struct(out)

table = Table.create


def map_arrays(data, f):
  if isinstance(data, np.ndarray): return f(data)
  if isinstance(data, Struct): return data.__class__({k: map_arrays(v, f) for k, v in data.items()})
  if isinstance(data, dict): return {k: map_arrays(v, f) for k, v in data.items()}
  if isinstance(data, (list, tuple)): return [map_arrays(v, f) for v in data]
  assert False, f"map_arrays: unsupported {type(data)}"


def reduce_arrays(data, f, op, initial=None):
  acc = initial
  def visit(d):
    nonlocal acc
    if isinstance(d, np.ndarray): acc = op(acc, f(d))
    elif isinstance(d, (Struct, dict)):
      for v in d.values(): visit(v)
    elif isinstance(d, (list, tuple)):
      for v in d: visit(v)
    else: assert False, f"reduce_arrays: unsupported {type(d)}"
  visit(data)
  return acc


def shape(data): return map_arrays(data, lambda a: a.shape)
def shape_info(data): return map_arrays(data, lambda a: (a.shape, a.dtype))
