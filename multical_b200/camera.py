"""Camera intrinsics as optimisable parameters, mirroring the reference's `Camera`
(multical/camera.py:27-181: model names 43-48, params 144-155, with_params 157-171) and
`CameraFisheye` (multical/camera_fisheye.py:28-170).  Projection itself runs on the GPU
(csrc/geometry.cuh); these classes only carry values across the API."""
from functools import cached_property

import numpy as np

from .parameters import Parameters
from .structs import struct

MODELS = ("standard", "rational", "thin_prism", "tilted")
DIST_SIZE = dict(standard=5, rational=8, thin_prism=12, tilted=14, fisheye=4)


class Camera(Parameters):
  engine_model = None      # resolved from .model

  def __init__(self, image_size, intrinsic, dist, model="standard", fix_aspect=False, has_skew=False, **_):
    assert model in MODELS, f"unknown camera model {model} options are {list(MODELS)}"
    self.model = model
    self.image_size = tuple(image_size)
    self.intrinsic = np.asarray(intrinsic, np.float64)
    self.dist = np.zeros(5) if dist is None else np.asarray(dist, np.float64)
    self.fix_aspect = fix_aspect
    self.has_skew = has_skew

  def __repr__(self): return f"Camera(intrinsic={self.intrinsic!r}, dist={self.dist!r}, image_size={self.image_size})"

  @property
  def focal_length(self): return np.array([self.intrinsic[0, 0], self.intrinsic[1, 1]])
  @property
  def principle_point(self): return np.array([self.intrinsic[0, 2], self.intrinsic[1, 2]])
  @property
  def skew(self): return self.intrinsic[0, 1] if self.has_skew else 0.0

  @cached_property
  def params(self):
    f = self.focal_length
    if self.fix_aspect: f = np.array([f.mean(), f.mean()])
    return struct(focal_length=f, principle_point=self.principle_point, skew=np.array([self.skew]), dist=self.dist)

  def with_params(self, params):
    f = params["focal_length"]
    fx, fy = (f[0], f[0]) if self.fix_aspect else (f[0], f[1])
    px, py = params["principle_point"]
    skew, = params["skew"]
    K = np.array([[fx, skew, px], [0, fy, py], [0, 0, 1]], np.float64)
    return self.copy(intrinsic=K, dist=params["dist"])

  def approx_eq(self, other):
    return self.image_size == other.image_size and np.allclose(other.intrinsic, self.intrinsic) and np.allclose(other.dist, self.dist)

  def scale_image(self, factor):
    K = self.intrinsic.copy(); K[:2] *= factor
    return self.copy(intrinsic=K)

  def __getstate__(self):
    return dict(image_size=self.image_size, intrinsic=self.intrinsic, dist=self.dist, fix_aspect=self.fix_aspect,
                has_skew=self.has_skew, model=self.model)
  def __setstate__(self, d): self.__dict__.update(d)
  def copy(self, **k):
    d = self.__getstate__(); d.update(k)
    return self.__class__(**d)


class CameraFisheye(Camera):
  def __init__(self, image_size, intrinsic, dist, model="standard", fix_aspect=False, has_skew=False, **_):
    self.model = model
    self.image_size = tuple(image_size)
    self.intrinsic = np.asarray(intrinsic, np.float64)
    self.dist = np.zeros((4, 1)) if dist is None else np.asarray(dist, np.float64)
    self.fix_aspect = fix_aspect
    self.has_skew = has_skew


def engine_model_of(camera):
  """Which CUDA camera model a (reference or mirror) camera object needs."""
  if type(camera).__name__ == "CameraFisheye": return "fisheye"
  model = getattr(camera, "model", "standard")
  nd = np.size(camera.dist)
  expect = DIST_SIZE[model]
  assert nd == expect, f"camera model {model} expects {expect} distortion coefficients, got {nd}"
  return model
