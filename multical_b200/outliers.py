"""Threshold selectors of the outlier loop and the order-statistic arithmetic of np.quantile.

The reference picks its rejection threshold and its robust-loss scale as `np.quantile(errors, q) * factor`
(select_threshold, multical/optimization/calibration.py:37-40; wired up from the command line in
config/arguments.py:59-63 and workspace.py:228-247).  With the point table resident on the GPU the error vector never
comes to the host: the engine sorts it there and hands back the two order statistics around each quantile; this module
holds the host half -- which ranks to ask for and how numpy interpolates between them (method='linear', the default)."""
import numpy as np


class QuantileThreshold:
  """errors -> np.quantile(errors, quantile) * factor, but inspectable, so that `Calibration.adjust_outliers` can evaluate
  it on the device.  Calling it on a host error vector gives exactly what the reference's closure gives."""

  def __init__(self, quantile=0.75, factor=5.0):
    assert 0.0 <= quantile <= 1.0, f"quantile {quantile} outside [0, 1]"
    self.quantile, self.factor = float(quantile), float(factor)

  def __call__(self, errors):
    return np.quantile(errors, self.quantile) * self.factor

  def __repr__(self):
    return f"QuantileThreshold(quantile={self.quantile}, factor={self.factor})"


def select_threshold(quantile=0.75, factor=5.0):
  """Same name, defaults and meaning as the reference's factory (calibration.py:37-40)."""
  return QuantileThreshold(quantile, factor)


def quantile_ranks(n, q):
  """The order statistics np.quantile(a, q) reads from sorted `a` of length n and its interpolation weight:
  (lower rank, upper rank, gamma), arrays shaped like q.  Follows numpy/lib/_function_base_impl.py (`_quantile`,
  `_QuantileMethods['linear']`, `_get_indexes`, `_get_gamma`) so the result is bit-identical."""
  q = np.asarray(q, dtype=np.float64)
  assert n > 0, "quantile of an empty error vector"
  virtual = (n - 1) * q               # the 'linear' entry of numpy's _QuantileMethods
  lower = np.floor(virtual)
  gamma = virtual - lower
  upper = lower + 1.0
  above = virtual >= n - 1
  lower = np.where(above, n - 1, lower)
  upper = np.where(above, n - 1, upper)
  below = virtual < 0
  lower = np.where(below, 0, lower)
  upper = np.where(below, 0, upper)
  return lower.astype(np.int64), upper.astype(np.int64), gamma


def lerp(a, b, t):
  """numpy's `_lerp`: a + (b-a)*t, evaluated from the other end for t >= 0.5."""
  a, b, t = np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(t, np.float64)
  d = b - a
  return np.where(t >= 0.5, b - d * (1.0 - t), a + d * t)


def quantile_from_sorted(fetch, n, q):
  """np.quantile(a, q) given only `fetch(ranks) -> sorted_a[ranks]` (the engine's mcba_table_error_ranks)."""
  lo, hi, gamma = quantile_ranks(n, q)
  ranks = np.concatenate([np.atleast_1d(lo), np.atleast_1d(hi)])
  vals = np.asarray(fetch(ranks), dtype=np.float64)
  k = ranks.size // 2
  out = lerp(vals[:k], vals[k:], np.atleast_1d(gamma))
  return out.reshape(np.shape(q)) if np.ndim(q) else float(out[0])
