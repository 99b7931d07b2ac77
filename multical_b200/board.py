"""Calibration target as the hot path sees it: a list of 3-D points (reference:
board.adjusted_points / num_points feeding tables.stack_boards, multical/tables.py:385-394;
params/with_params: board/charuco.py:112-117).  Detection / drawing are out of scope."""
from functools import cached_property

import numpy as np

from .parameters import Parameters


class Board(Parameters):
  """size / min_points / min_rows / id_divisor: the id grid and thresholds of `has_min_detections` (board/common.py:30-34;
  CharucoBoard: corner ids on the (w, h) = size grid, min_rows=3, min_points=20, charuco.py:12,104-106; AprilGrid: tag id =
  corner id // 4, aprilgrid.py:197-199) -- only used by the batched pose initialisation (multical_b200/tables.py)."""
  def __init__(self, adjusted_points, size=None, min_points=20, min_rows=3, id_divisor=1):
    self.adjusted_points = np.asarray(adjusted_points)
    self.size, self.min_points, self.min_rows, self.id_divisor = size, min_points, min_rows, id_divisor

  @property
  def points(self): return self.adjusted_points
  @property
  def num_points(self): return self.adjusted_points.shape[0]
  @cached_property
  def params(self): return self.adjusted_points
  def with_params(self, params): return Board(params, self.size, self.min_points, self.min_rows, self.id_divisor)
  def has_min_detections(self, detections):
    ids = np.asarray(detections.ids) // self.id_divisor
    w, h = self.size
    rows, cols = np.unravel_index(ids, (h, w))
    return ids.size >= self.min_points and np.unique(rows).size >= self.min_rows and np.unique(cols).size >= self.min_rows
  def __getstate__(self):
    return dict(adjusted_points=self.adjusted_points, size=self.size, min_points=self.min_points, min_rows=self.min_rows, id_divisor=self.id_divisor)
  def __setstate__(self, d): self.__dict__.update(d)


def stack_boards(boards):
  """Pad every board to the largest point count (tables.py:385-394)."""
  P = max(b.num_points for b in boards)
  pts = np.zeros((len(boards), P, 3)); valid = np.zeros((len(boards), P), bool)
  for i, b in enumerate(boards):
    n = b.num_points
    pts[i, :n] = np.asarray(b.adjusted_points, dtype=np.float64)
    valid[i, :n] = True
  return pts, valid
