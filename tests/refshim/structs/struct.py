"""Minimal insertion-ordered attribute dict with the py-structs method names the
reference hot path touches (see tests/refshim/structs/__init__.py)."""
from collections.abc import Mapping


class Struct(Mapping):
  def __init__(self, entries=None, **kw):
    d = dict(entries or {})
    d.update(kw)
    self.__dict__.update(d)

  # Mapping protocol
  def __getitem__(self, k): return self.__dict__[k]
  def __iter__(self): return iter(self.__dict__)
  def __len__(self): return len(self.__dict__)
  def __contains__(self, k): return k in self.__dict__
  def keys(self): return self.__dict__.keys()
  def values(self): return self.__dict__.values()
  def items(self): return self.__dict__.items()
  def get(self, k, default=None): return self.__dict__.get(k, default)
  def __setitem__(self, k, v): self.__dict__[k] = v
  def __repr__(self):
    return "struct(" + ", ".join(f"{k}={v!r}" for k, v in self.items()) + ")"
  def __eq__(self, other):
    return isinstance(other, Struct) and self.__dict__ == other.__dict__
  def __hash__(self): return id(self)
  def __getstate__(self): return dict(self.__dict__)
  def __setstate__(self, d): self.__dict__.update(d)

  def _to_dicts(self): return to_dicts(self)
  def _subset(self, *keys): return self.__class__({k: self[k] for k in keys})
  def _without(self, *keys):
    return self.__class__({k: v for k, v in self.items() if k not in keys})
  def _filter(self, f): return self.__class__({k: v for k, v in self.items() if f(v)})
  def _filterWithKey(self, f):
    return self.__class__({k: v for k, v in self.items() if f(k)})
  def _map(self, f, *args, **kw):
    return self.__class__({k: f(v, *args, **kw) for k, v in self.items()})
  def _mapWithKey(self, f):
    return self.__class__({k: f(k, v) for k, v in self.items()})
  def _zipWith(self, f, *others):
    assert all(o.keys() == self.keys() for o in others)
    return self.__class__({k: f(v, *[o[k] for o in others]) for k, v in self.items()})
  def _extend(self, **extra):
    d = dict(self.__dict__); d.update(extra)
    return self.__class__(d)
  def _update(self, **extra):
    for k in extra: assert k in self.__dict__, f"_update: unknown key {k}"
    return self._extend(**extra)
  def _merge(self, other):
    d = dict(self.__dict__); d.update(dict(other.items()))
    return self.__class__(d)


def struct(**d): return Struct(d)
def to_structs(d):
  if isinstance(d, dict): return Struct({k: to_structs(v) for k, v in d.items()})
  if isinstance(d, list): return [to_structs(v) for v in d]
  return d
def to_dicts(s):
  if isinstance(s, Struct): return {k: to_dicts(v) for k, v in s.items()}
  if isinstance(s, dict): return {k: to_dicts(v) for k, v in s.items()}
  if isinstance(s, list): return [to_dicts(v) for v in s]
  return s
def subset(d, keys): return {k: d[k] for k in keys}
def choose(*options):
  for o in options:
    if o is not None: return o
  assert False, "choose: all options were None"
def when(cond, x): return x if cond else None
def apply_none(f, x): return None if f is None else f(x)
def map_none(f, x): return None if x is None else f(x)
def concat_lists(xs): return [x for inner in xs for x in inner]
def map_list(f, xs, **kw): return [f(x, **kw) for x in xs]
def filter_none(xs): return [x for x in xs if x is not None]
def split_list(xs, sizes):
  out, i = [], 0
  for n in sizes:
    out.append(xs[i:i + n]); i += n
  return out
def split_dict(d): return list(d.keys()), list(d.values())
def transpose_lists(lists): return list(map(list, zip(*lists)))
def transpose_structs(structs):
  elem = structs[0]
  return elem.__class__({k: [s[k] for s in structs] for k in elem.keys()})
def invert_keys(d):
  """{k: v} -> {v: k} (used as tables.dimension_name, tables.py:97)."""
  return d.__class__({v: k for k, v in d.items()}) if isinstance(d, Struct) else {v: k for k, v in d.items()}
