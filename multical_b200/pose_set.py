"""Named set of 4x4 poses + valid flags whose parameters are rtvecs
(reference: multical/optimization/pose_set.py:12-73; StaticFrames: motion/static_frames.py:29-42)."""
from functools import cached_property

import numpy as np

from . import rtvec
from .parameters import Parameters
from .structs import Table


class PoseSet(Parameters):
  def __init__(self, pose_table, names=None):
    self.pose_table = pose_table
    self.names = names or [str(i) for i in range(self.size)]

  @property
  def size(self): return self.poses.shape[0]
  @property
  def valid(self): return self.pose_table.valid
  @property
  def poses(self): return self.pose_table.poses

  def __getitem__(self, k):
    if isinstance(k, str):
      if k not in self.names: raise KeyError(f"pose {k} not found in {self.names}")
      return self.poses[self.names.index(k)]
    return self.poses[k]

  def relative(self, src, dest): return self[dest] @ np.linalg.inv(self[src])
  def pre_transform(self, t): return self.copy(pose_table=self.pose_table._extend(poses=t @ self.poses))
  def post_transform(self, t): return self.copy(pose_table=self.pose_table._extend(poses=self.poses @ t))

  @cached_property
  def params(self): return rtvec.from_matrix(self.poses).ravel()
  def with_params(self, params):
    return self.copy(pose_table=self.pose_table._update(poses=rtvec.to_matrix(np.asarray(params).reshape(-1, 6))))

  def __getstate__(self): return dict(pose_table=self.pose_table, names=self.names)
  def __setstate__(self, d): self.__dict__.update(d)
  def copy(self, **k):
    d = self.__getstate__(); d.update(k)
    return self.__class__(**d)


class StaticFrames(PoseSet):
  """One rig pose per frame (the only motion model on the BASELINE configs)."""
  def __init__(self, pose_table, names=None):
    super().__init__(pose_table, names)

  @staticmethod
  def init(pose_table, names=None): return StaticFrames(pose_table, names)
  @property
  def frame_poses(self): return self.pose_table


def pose_table(poses, valid=None):
  poses = np.asarray(poses, np.float64)
  valid = np.ones(poses.shape[0], bool) if valid is None else np.asarray(valid, bool)
  return Table.create(poses=poses, valid=valid)
