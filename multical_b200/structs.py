"""Tiny attribute-dict / struct-of-arrays helpers for the host side.

The reference leans on the third-party `py-structs` package for this plumbing
(e.g. calibration.py:21-22, parameters.py:6); it is absent here, so the host mirror
carries its own minimal equivalents (no arithmetic lives here)."""
import numpy as np


class Struct(dict):
  """dict with attribute access; insertion ordered (parameter order depends on it)."""
  __getattr__ = dict.__getitem__

  def __setattr__(self, k, v): self[k] = v
  def __getstate__(self): return dict(self)
  def __setstate__(self, d): self.update(d)
  def _extend(self, **kw):
    out = self.__class__(self); out.update(kw); return out
  def _update(self, **kw):
    for k in kw: assert k in self, f"unknown key {k}"
    return self._extend(**kw)
  def _map(self, f): return self.__class__({k: f(v) for k, v in self.items()})


def struct(**kw): return Struct(kw)


class Table(Struct):
  """Struct of numpy arrays sharing a leading shape (`_prefix`)."""
  @staticmethod
  def create(**arrays): return Table({k: np.asarray(v) for k, v in arrays.items()})

  @property
  def _prefix(self):
    shapes = [np.shape(v) for v in self.values()]
    out = []
    for dims in zip(*shapes):
      if all(d == dims[0] for d in dims): out.append(dims[0])
      else: break
    return tuple(out)
  _shape = _prefix
