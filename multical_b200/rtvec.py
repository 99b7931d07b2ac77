"""Host-side pose encoding helpers: [rx ry rz tx ty tz] <-> 4x4 (reference: transform/rtvec.py:16-32,
transform/matrix.py:33-44).  Only used to move poses across the API; the solver's own Rodrigues /
Jacobians live in csrc/geometry.cuh."""
import numpy as np
from scipy.spatial.transform import Rotation

size = 6


def to_matrix(rtvec):
  rtvec = np.asarray(rtvec, np.float64)
  lead = rtvec.shape[:-1]
  flat = rtvec.reshape(-1, 6)
  m = np.tile(np.eye(4), (flat.shape[0], 1, 1))
  m[:, :3, :3] = Rotation.from_rotvec(flat[:, :3]).as_matrix()
  m[:, :3, 3] = flat[:, 3:]
  return m.reshape(*lead, 4, 4)


def from_matrix(m):
  m = np.asarray(m, np.float64)
  lead = m.shape[:-2]
  flat = m.reshape(-1, 4, 4)
  out = np.hstack([Rotation.from_matrix(flat[:, :3, :3]).as_rotvec(), flat[:, :3, 3]])
  return out.reshape(*lead, 6)
