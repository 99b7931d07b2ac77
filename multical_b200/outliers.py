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


def merged_order_statistics(comm, fetch_local, count_below_local, n_local, ranks, splitters=256):
  """Order statistics of the UNION of every rank's error vector, exactly, without moving the vectors: `ranks` (0-based, global) ->
  values, identical on all ranks.  Every rank holds its own errors sorted on its GPU and answers two questions about them:
  `fetch_local(local_ranks) -> values` (mcba_table_error_ranks) and `count_below_local(values) -> counts` (mcba_table_count_below).
    1. all-gather `splitters` evenly spaced local order statistics per rank -> a sorted candidate set V that contains the global max;
    2. all-reduce the local counts below every candidate -> C(v) = global number of errors < v; the k-th error lies in the gap
       [v_lo, v_hi) around C(v_lo) <= k < C(v_hi), which holds only ~N / splitters errors;
    3. all-gather the local errors of those gaps (consecutive local ranks c_r(v_lo) .. c_r(v_hi)), merge, index.
  `comm`: all_gather(obj) -> list over ranks, all_reduce_sum(int64 array) -> array."""
  ranks = np.atleast_1d(np.asarray(ranks, dtype=np.int64))
  # 1. candidates
  if n_local > 0:
    pick = np.unique(np.round(np.linspace(0, n_local - 1, min(n_local, splitters))).astype(np.int64))
    mine = np.asarray(fetch_local(pick), dtype=np.float64)
  else:
    mine = np.zeros(0)
  V = np.unique(np.concatenate(comm.all_gather(mine)))
  assert V.size > 0, "order statistic of an empty error vector"
  # 2. global counts below every candidate (and the local ones: they delimit the local share of every gap)
  c_local = np.asarray(count_below_local(V), dtype=np.int64) if n_local > 0 else np.zeros(V.size, np.int64)
  C = comm.all_reduce_sum(c_local)
  total = int(comm.all_reduce_sum(np.array([n_local], np.int64))[0])
  assert ranks.min() >= 0 and ranks.max() < total, "error rank out of range"
  lo_idx = np.searchsorted(C, ranks, side="right") - 1            # last candidate with C(v) <= k   (C[0] = 0: V[0] is the global min)
  # 3. the local errors of each needed gap [V[lo], V[lo+1])  (the last gap is everything >= the global maximum, i.e. copies of it)
  gaps = np.unique(lo_idx)
  local_lo = c_local[gaps]
  local_hi = np.where(gaps + 1 < V.size, c_local[np.minimum(gaps + 1, V.size - 1)], n_local)
  need = np.concatenate([np.arange(a, b) for a, b in zip(local_lo, local_hi)]) if gaps.size else np.zeros(0, np.int64)
  vals = np.asarray(fetch_local(need), dtype=np.float64) if need.size else np.zeros(0)
  sizes = (local_hi - local_lo).astype(np.int64)
  parts = comm.all_gather((sizes, vals))
  out = np.zeros(ranks.size)
  for gi, g in enumerate(gaps):
    chunks = []
    for sz, v in parts:
      off = int(sz[:gi].sum())
      chunks.append(v[off:off + int(sz[gi])])
    merged = np.sort(np.concatenate(chunks))
    sel = lo_idx == g
    out[sel] = merged[ranks[sel] - C[g]]
  return out
