"""Pose sets on the host side of the boundary.

A pose set is N rigid transforms (4x4, "points are mapped": x' = T x) with a validity flag each and an optional name
each; as optimisation parameters it is the flat vector of N rtvecs [rx ry rz tx ty tz] -- every pose, valid or not, in
index order.  That convention is what fixes the layout of the camera-pose, board-pose and motion blocks of the parameter
vector, so it mirrors the reference exactly (multical/optimization/pose_set.py:12-73, params 51-53, with_params 55-57;
`StaticFrames` = one rig pose per frame, multical/motion/static_frames.py:29-42).  No projection arithmetic lives here: the
solver consumes the rtvecs through `Calibration._state_arrays` and works on the GPU.
"""
from functools import cached_property

import numpy as np

from . import rtvec
from .parameters import Parameters
from .structs import Table


def pose_table(poses, valid=None):
  """Table(poses f64[N,4,4], valid bool[N]) from plain arrays (all valid when `valid` is omitted)."""
  poses = np.asarray(poses, dtype=np.float64)
  flags = np.ones(len(poses), dtype=bool) if valid is None else np.asarray(valid, dtype=bool)
  assert poses.shape == (len(flags), 4, 4), f"expected [N,4,4] poses, got {poses.shape}"
  return Table.create(poses=poses, valid=flags)


class PoseSet(Parameters):
  """N named poses.  Immutable in use: every modifier returns a new set built by `copy`."""

  _state_keys = ("pose_table", "names")

  def __init__(self, pose_table, names=None):
    self.pose_table = pose_table
    count = np.shape(pose_table.poses)[0]
    self.names = list(names) if names else [str(i) for i in range(count)]

  # -- plain accessors --------------------------------------------------------------------------
  poses = property(lambda self: self.pose_table.poses)
  valid = property(lambda self: self.pose_table.valid)
  size = property(lambda self: np.shape(self.pose_table.poses)[0])

  def index_of(self, key):
    if not isinstance(key, str):
      return key
    try:
      return self.names.index(key)
    except ValueError:
      raise KeyError(f"pose {key} not found in {self.names}") from None

  def __getitem__(self, key):
    return self.poses[self.index_of(key)]

  def relative(self, src, dest):
    """Transform taking pose `src` to pose `dest`."""
    return self[dest] @ np.linalg.inv(self[src])

  # -- re-referencing (used by Calibration.transform_views / with_master) -----------------------
  def _with_poses(self, new_poses):
    return self.copy(pose_table=self.pose_table._extend(poses=new_poses))

  def pre_transform(self, t):
    return self._with_poses(np.asarray(t) @ self.poses)

  def post_transform(self, t):
    return self._with_poses(self.poses @ np.asarray(t))

  @cached_property
  def inverse(self):
    return self._with_poses(np.linalg.inv(self.poses))

  # -- parameters ---------------------------------------------------------------------------------
  @cached_property
  def params(self):
    return rtvec.from_matrix(self.poses).reshape(-1)

  def with_params(self, params):
    rt = np.asarray(params, dtype=np.float64).reshape(self.size, rtvec.size)
    return self.copy(pose_table=self.pose_table._update(poses=rtvec.to_matrix(rt)))

  @property
  def num_params(self):
    return rtvec.size * self.size

  def with_param_vec(self, param_vec):
    # the generic path would convert every pose to an rtvec just to learn the shape (6 per pose) of the vector to split
    param_vec = np.asarray(param_vec)
    assert param_vec.size == self.num_params, f"inconsistent parameter sizes, got {param_vec.size}, expected {self.num_params}"
    return self.with_params(param_vec)

  # -- copy / pickle -------------------------------------------------------------------------------
  def __getstate__(self):
    return {k: getattr(self, k) for k in self._state_keys}

  def __setstate__(self, state):
    self.__dict__.update(state)

  def copy(self, **changes):
    state = self.__getstate__()
    state.update(changes)
    return type(self)(**state)


class StaticFrames(PoseSet):
  """Motion model of the BASELINE configurations: the rig has one pose per frame, so the `motion` parameter block is a
  pose set over frames and a corner's pose chain is T_cam[c] T_frame[f] T_board[b]."""

  def __init__(self, pose_table, names=None):
    super().__init__(pose_table, names)

  @staticmethod
  def init(pose_table, names=None):
    return StaticFrames(pose_table, names)

  @property
  def frame_poses(self):
    return self.pose_table
