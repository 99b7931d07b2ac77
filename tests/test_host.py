"""CPU: host-side logic (parameter plumbing, packing order, C-ABI surface) -- no compute calls."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_CASES, ROOT, load_golden, optimize_of
from multical_b200 import _native, parameters, rtvec, synthetic
from multical_b200.calibration import Calibration, default_optimize, error_stats, from_scene, select_threshold
from multical_b200.engine import format_log, pack_corners
from oracle.ba_oracle import Problem


def test_cabi_library_exports_every_declared_symbol(build_lib):
  header = open(os.path.join(ROOT, "include", "mcba.h")).read()
  declared = sorted(set(re.findall(r"\b(mcba_[a-z0-9_]+)\s*\(", header)))
  assert len(declared) >= 15
  lib = _native.load()
  for sym in declared:
    assert hasattr(lib, sym), f"libmcba.so does not export {sym}"
  assert sorted(_native.EXPORTS) == declared
  assert lib.mcba_version() >= 100


def test_engine_fails_loudly_without_gpu(build_lib):
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from multical_b200.engine import Engine
  with pytest.raises(_native.NativeError):
    Engine(0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_param_vec_layout_matches_reference(name):
  """Calibration.param_vec must equal the reference's vector bit for bit (parameters.py:104-106,
  calibration.py:146-161, camera.py:150-155, pose_set.py:51-53)."""
  scene, z = load_golden(name)
  calib = from_scene(scene)
  if bool(z["cameras_enabled"]): calib = calib.enable(cameras=True)
  assert np.array_equal(calib.param_vec, z["x0"])
  # round trip through with_param_vec
  # (the reference drops a perturbed skew back to 0 unless has_skew: camera.py:139-141,163-171)
  again = calib.with_param_vec(z["x1"])
  expect = Problem.from_scene(scene, optimize=optimize_of(z)).with_param_vec(z["x1"]).param_vec
  assert np.allclose(again.param_vec, expect, rtol=0, atol=1e-12)
  S = calib.sparsity_matrix.tocsr(); S.sort_indices()
  assert np.array_equal(S.indptr, z["sp_indptr"]) and np.array_equal(S.indices, z["sp_indices"])


def test_pack_corners_order_is_boolean_mask_order():
  scene = synthetic.make_scene(C=2, F=3, vis=0.4, seed=2)
  idx, obs = pack_corners(scene["valid"], scene["points"])
  assert np.array_equal(idx, np.argwhere(scene["valid"]))
  assert np.array_equal(obs, scene["points"][scene["valid"]])
  lin = np.ravel_multi_index(idx.T, scene["valid"].shape)
  assert np.all(np.diff(lin) > 0)           # row-major, strictly increasing


def test_parameters_split_join_roundtrip():
  tree = dict(a=np.arange(6.0).reshape(2, 3), b=[np.arange(4.0), np.arange(2.0)])
  v = parameters.join(tree)
  assert v.tolist() == [0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 0, 1]
  back = parameters.split(v * 2, tree)
  assert back["a"].shape == (2, 3) and back["b"][1].tolist() == [0, 2]
  with pytest.raises(AssertionError):
    parameters.split(v[:-1], tree)


def test_rtvec_roundtrip():
  rng = np.random.default_rng(0)
  rt = rng.normal(0, 0.7, (20, 6))
  assert np.allclose(rtvec.from_matrix(rtvec.to_matrix(rt)), rt, atol=1e-12)


def test_calibration_api_surface():
  scene = synthetic.make_scene(C=2, F=3, vis=0.4, seed=2)
  calib = from_scene(scene)
  assert dict(calib.size) == dict(cameras=2, rig_poses=3, boards=1, points=315)
  assert calib.optimize == default_optimize and calib.param_vec.size == 6 * (2 + 1 + 3)
  with pytest.raises(AssertionError):
    calib.enable(bogus=True)
  c2 = calib.enable(cameras=True)
  assert c2.param_vec.size == 6 * 6 + 2 * 10 and calib.optimize["cameras"] is False
  import pickle
  c3 = pickle.loads(pickle.dumps(c2))
  assert np.array_equal(c3.param_vec, c2.param_vec)
  assert sorted(c2.__getstate__()) == sorted(["cameras", "boards", "point_table", "camera_poses", "board_poses", "motion", "inlier_mask", "optimize"])
  # with_master keeps the product T_cam T_frame unchanged
  m = c2.with_master(1)
  a = c2.camera_poses.poses[:, None] @ c2.motion.poses[None]
  b = m.camera_poses.poses[:, None] @ m.motion.poses[None]
  assert np.allclose(a, b, atol=1e-12) and np.allclose(m.camera_poses.poses[1], np.eye(4), atol=1e-12)
  assert np.isclose(select_threshold(0.5, 2.0)(np.arange(5.0)), 4.0)
  st = error_stats(np.array([3.0, 4.0]))
  assert np.isclose(st.rms, np.sqrt(12.5)) and st.n == 2
  assert error_stats(np.zeros(0)).n == 1         # empty-array guard of the reference (calibration.py:304-306)


def test_valid_mask_matches_oracle():
  scene, z = load_golden("invalid_poses_3x6")
  calib = from_scene(scene)
  prob = Problem.from_scene(scene)
  assert np.array_equal(calib.valid, prob.valid)
  assert calib.valid.sum() == z["r0"].size // 2


def test_log_table_format():
  lines = format_log([(0, 1, 2.8654e6, float("nan"), float("nan"), 9.18e7), (1, 2, 784.08, 2.86e6, 11.9, 3.22e5)])
  assert lines[0].split() == ["Iteration", "Total", "nfev", "Cost", "Cost", "reduction", "Step", "norm", "Optimality"]
  assert lines[1].split() == ["0", "1", "2.8654e+06", "9.18e+07"]
  assert lines[2].split() == ["1", "2", "7.8408e+02", "2.86e+06", "1.19e+01", "3.22e+05"]


def _run_outlier_loop(monkeypatch, scene, host, **kwargs):
  """adjust_outliers over the numpy stand-in engine (tests/fake_engine.py); returns (result, log lines, engine calls)."""
  import logging
  from fake_engine import FakeEngine
  from multical_b200 import calibration
  eng = FakeEngine()
  monkeypatch.setattr(calibration, "get_engine", lambda device=None: eng)
  if host: monkeypatch.setenv("MCBA_HOST_OUTLIERS", "1")
  else: monkeypatch.delenv("MCBA_HOST_OUTLIERS", raising=False)
  lines = []
  handler = logging.Handler(); handler.emit = lambda rec: lines.append(rec.getMessage())
  logger = logging.getLogger("calibration"); old = logger.level
  logger.addHandler(handler); logger.setLevel(logging.INFO)
  try:
    out = from_scene(scene).enable(cameras=True).adjust_outliers(**kwargs)
  finally:
    logger.removeHandler(handler); logger.setLevel(old)
  return out, lines, eng.calls


@pytest.mark.parametrize("with_scale", [False, True])
def test_resident_outlier_loop_is_the_host_loop(monkeypatch, with_scale):
  """The device-resident loop must make the decisions of the host loop (calibration.py:250-266): same thresholds, same
  masks, same log lines -- checked over a numpy engine so that only the control flow is under test here."""
  scene = synthetic.make_scene(C=2, F=5, vis=0.6, seed=43, outlier_fraction=0.03)
  kw = dict(num_adjustments=2, select_outliers=select_threshold(quantile=0.75, factor=5.0), tolerance=1e-6, max_iterations=30)
  if with_scale: kw.update(select_scale=select_threshold(quantile=0.5, factor=4.0), loss="soft_l1")
  host, host_log, host_calls = _run_outlier_loop(monkeypatch, scene, host=True, **kw)
  res, res_log, res_calls = _run_outlier_loop(monkeypatch, scene, host=False, **kw)
  assert res_calls.count("table_upload") == 1 and "upload_dense" not in res_calls       # the table crosses once
  assert host_calls.count("upload_dense") >= 5
  assert [c for c in res_calls if c.startswith("solve")] == [c for c in host_calls if c.startswith("solve")]
  assert res_log == host_log
  assert any(l.startswith("Rejecting") for l in res_log) and sum(l.startswith("Adjust_outliers") for l in res_log) == 3
  assert np.array_equal(res.inlier_mask, host.inlier_mask) and res.inlier_mask.sum() < res.valid.sum()
  assert np.allclose(res.camera_poses.poses, host.camera_poses.poses, atol=1e-12)
  assert np.allclose(res.cameras.param_vec, host.cameras.param_vec, atol=1e-12)


def test_resident_loop_falls_back_for_opaque_selectors(monkeypatch):
  """A plain callable cannot be evaluated on the device: the loop must then take the host path, not guess."""
  scene = synthetic.make_scene(C=2, F=4, vis=0.6, seed=44, outlier_fraction=0.03)
  rule = lambda errors: np.quantile(errors, 0.75) * 5.0
  _, _, calls = _run_outlier_loop(monkeypatch, scene, host=False, num_adjustments=1, select_outliers=rule, max_iterations=10)
  assert "table_upload" not in calls and "upload_dense" in calls


def test_quantile_from_order_statistics_is_numpy_quantile():
  from multical_b200.outliers import quantile_from_sorted
  rng = np.random.default_rng(5)
  for trial in range(300):
    n = int(rng.integers(1, 40)) if trial % 2 else int(rng.integers(1, 50000))
    a = np.abs(rng.standard_normal(n)) * rng.uniform(0.1, 10)
    srt = np.sort(a)
    q = np.append(rng.uniform(0, 1, 4), rng.choice([0.0, 0.25, 0.5, 0.75, 0.95, 1.0]))
    assert np.array_equal(quantile_from_sorted(lambda r: srt[r], n, q), np.quantile(a, q))
    assert quantile_from_sorted(lambda r: srt[r], n, 0.75) == np.quantile(a, 0.75)
  thr = select_threshold(0.75, 5.0)
  assert thr(srt) == np.quantile(srt, 0.75) * 5.0 and (thr.quantile, thr.factor) == (0.75, 5.0)


def test_outlier_steps_follow_reference_golden_over_numpy_engine(monkeypatch):
  """Host plumbing of reject_outliers / reject_outliers_quantile / adjust_outliers (mask order, thresholds, resident loop)
  against what the running reference produced (tests/golden/outliers_3x6.npz); errors come from the numpy engine."""
  from fake_engine import FakeEngine
  from multical_b200 import calibration
  scene, z = load_golden("outliers_3x6")
  eng = FakeEngine()
  monkeypatch.setattr(calibration, "get_engine", lambda device=None: eng)
  monkeypatch.delenv("MCBA_HOST_OUTLIERS", raising=False)
  calib = from_scene(scene).enable(cameras=True)
  assert np.abs(calib.reprojection_error - z["err_valid"]).max() < 1e-9
  thr = select_threshold(quantile=0.75, factor=5.0)(calib.reprojection_error)
  assert abs(thr - float(z["thr_q75x5"])) < 1e-9
  assert np.array_equal(calib.reject_outliers(float(z["thr_q75x5"])).inliers, z["inliers_thr"])
  assert np.array_equal(calib.reject_outliers_quantile(0.95).inliers, z["inliers_q95"])
  out = calib.adjust_outliers(num_adjustments=2, select_outliers=select_threshold(quantile=0.75, factor=5.0))
  assert "table_upload" in eng.calls
  assert np.array_equal(out.inlier_mask, z["adj_inliers"])
  assert abs(np.sqrt(np.mean(out.reprojection_inliers ** 2)) - float(z["adj_rms"])) < 2e-3
