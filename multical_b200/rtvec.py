"""Host-side pose encoding helpers: [rx ry rz tx ty tz] <-> 4x4 (reference: transform/rtvec.py:16-32,
transform/matrix.py:33-44).  Only used to move poses across the API; the solver's own Rodrigues /
Jacobians live in csrc/geometry.cuh."""
import numpy as np
from scipy.spatial.transform import Rotation

size = 6


def to_matrix(rtvec):
  """[..., 6] rtvecs -> [..., 4, 4].  Rodrigues formula in plain numpy (a scipy Rotation round trip costs ~70 us per call,
  which is visible next to a 0.9 ms solve); agrees with scipy's from_rotvec(...).as_matrix() to 1e-15."""
  rtvec = np.asarray(rtvec, np.float64)
  lead = rtvec.shape[:-1]
  flat = rtvec.reshape(-1, 6)
  r = flat[:, :3]
  th2 = np.einsum("ij,ij->i", r, r)
  small = th2 < 1e-8
  th = np.sqrt(np.where(small, 1.0, th2))
  a = np.where(small, 1.0 - th2 / 6.0, np.sin(th) / th)                      # sin(t)/t
  b = np.where(small, 0.5 - th2 / 24.0, (1.0 - np.cos(th)) / np.where(small, 1.0, th2))   # (1-cos t)/t^2
  x, y, z = r[:, 0], r[:, 1], r[:, 2]
  m = np.zeros((flat.shape[0], 4, 4))
  m[:, 0, 0] = 1.0 - b * (y * y + z * z); m[:, 0, 1] = b * x * y - a * z;       m[:, 0, 2] = b * x * z + a * y
  m[:, 1, 0] = b * x * y + a * z;       m[:, 1, 1] = 1.0 - b * (x * x + z * z); m[:, 1, 2] = b * y * z - a * x
  m[:, 2, 0] = b * x * z - a * y;       m[:, 2, 1] = b * y * z + a * x;       m[:, 2, 2] = 1.0 - b * (x * x + y * y)
  m[:, :3, 3] = flat[:, 3:]
  m[:, 3, 3] = 1.0
  return m.reshape(*lead, 4, 4)


def from_matrix(m):
  m = np.asarray(m, np.float64)
  lead = m.shape[:-2]
  flat = m.reshape(-1, 4, 4)
  out = np.hstack([Rotation.from_matrix(flat[:, :3, :3]).as_rotvec(), flat[:, :3, 3]])
  return out.reshape(*lead, 6)
