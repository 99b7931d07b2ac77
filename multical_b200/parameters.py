"""Parameter-tree <-> flat vector plumbing with the reference's names and order
(multical/optimization/parameters.py:28-106).  Trees are nested dict/Struct/list of arrays,
flattened depth-first in insertion order."""
from functools import cached_property

import numpy as np


def _leaves(tree, out):
  if isinstance(tree, np.ndarray): out.append(tree)
  elif isinstance(tree, dict):
    for v in tree.values(): _leaves(v, out)
  elif isinstance(tree, (list, tuple)):
    for v in tree: _leaves(v, out)
  else: raise TypeError(f"unsupported parameter node {type(tree)}")
  return out


def count(params): return sum(a.size for a in _leaves(params, []))


def join(params):
  leaves = _leaves(params, [])
  return np.concatenate([a.ravel() for a in leaves]) if leaves else np.zeros(0)


def split(param_vec, params):
  total = count(params)
  assert param_vec.size == total, f"inconsistent parameter sizes, got {param_vec.size}, expected {total}"
  pos = 0
  def take(tree):
    nonlocal pos
    if isinstance(tree, np.ndarray):
      out = param_vec[pos:pos + tree.size].reshape(tree.shape); pos += tree.size
      return out
    if isinstance(tree, dict): return tree.__class__({k: take(v) for k, v in tree.items()})
    return [take(v) for v in tree]
  return take(params)


class Parameters:
  @cached_property
  def params(self): raise NotImplementedError()
  def with_params(self, params): raise NotImplementedError()
  @cached_property
  def param_vec(self): return join(self.params)
  def with_param_vec(self, param_vec): return self.with_params(split(np.asarray(param_vec), self.params))


class ParamList(Parameters):
  def __init__(self, param_objects, names=None):
    self.param_objects = list(param_objects)
    self.names = names

  def __getitem__(self, index):
    if isinstance(index, str) and self.names is not None: index = self.names.index(index)
    return self.param_objects[index]
  def __iter__(self): return iter(self.param_objects)
  def __len__(self): return len(self.param_objects)
  def __repr__(self): return f"ParamList({self.param_objects!r})"

  @cached_property
  def params(self): return [p.param_vec for p in self.param_objects]
  @property
  def num_params(self): return sum(p.param_vec.size for p in self.param_objects)
  def with_params(self, params):
    return ParamList([o.with_param_vec(p) for o, p in zip(self.param_objects, params)], self.names)
