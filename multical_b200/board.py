"""Calibration target as the hot path sees it: a list of 3-D points (reference:
board.adjusted_points / num_points feeding tables.stack_boards, multical/tables.py:385-394;
params/with_params: board/charuco.py:112-117).  Detection / drawing are out of scope."""
from functools import cached_property

import numpy as np

from .parameters import Parameters


class Board(Parameters):
  def __init__(self, adjusted_points):
    self.adjusted_points = np.asarray(adjusted_points)

  @property
  def points(self): return self.adjusted_points
  @property
  def num_points(self): return self.adjusted_points.shape[0]
  @cached_property
  def params(self): return self.adjusted_points
  def with_params(self, params): return Board(params)
  def __getstate__(self): return dict(adjusted_points=self.adjusted_points)
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
