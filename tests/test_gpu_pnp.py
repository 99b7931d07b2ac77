"""GPU (-m gpu): batched board-pose initialisation (include/mcba.h mcba_pnp_views, csrc/pnp_kernels.cuh)
against the outputs of the reference's own tables.make_pose_table (tests/golden/pnp_*.npz) and against OpenCV through the oracle.

Bars: validity flags and corner counts exact; pose (4x4, board wrt camera) <= 1e-6 absolute (observed ~1e-9: OpenCV's own LM ends
within ~3e-10 of the minimiser the device iterates to); reprojection RMS <= 1e-6 px (OpenCV reports it in float32: ~1.5e-8);
view angles <= 1e-4 degrees."""
import numpy as np
import pytest

from conftest import GOLDEN
from multical_b200 import synthetic
from multical_b200.board import Board
from multical_b200.calibration import from_scene, get_engine
from multical_b200.camera import Camera, CameraFisheye
from multical_b200.structs import Table
from multical_b200.tables import detection_lists, make_pose_table
from oracle import pnp_oracle

pytestmark = pytest.mark.gpu
PNP_CASES = ["pnp_std_3x6", "pnp_fisheye_2x5", "pnp_cube_3x4"]


def load(name):
  z = dict(np.load(f"{GOLDEN}/{name}.npz", allow_pickle=False))
  model, size = str(z["model"]), tuple(int(v) for v in z["image_size"])
  cams = [(CameraFisheye(size, K, d) if model == "fisheye" else Camera(size, K, d, model=model)) for K, d in zip(z["K"], z["dist"])]
  w, h, div, min_points, min_rows = (int(v) for v in z["grid"])
  boards = [Board(p, size=(w, h), min_points=min_points, min_rows=min_rows, id_divisor=div) for p in z["board_points"]]
  return z, Table.create(points=z["points"], valid=z["valid"]), boards, cams


@pytest.mark.parametrize("name", PNP_CASES)
def test_pose_table_matches_reference_golden(name):
  z, table, boards, cams = load(name)
  got = make_pose_table(table, boards, cams, bool(z["exclude_bad_poses"]), float(z["pose_error_limit"]))
  ok = z["pose_valid"]
  assert 0 < ok.sum() < ok.size or name == "pnp_fisheye_2x5"              # the fixtures contain rejected views
  assert np.array_equal(np.asarray(got.valid), ok)
  assert np.array_equal(np.asarray(got.num_points), z["num_points"])
  assert np.abs(np.asarray(got.poses) - z["poses"]).max() < 1e-6          # invalid views: identity on both sides
  assert np.abs(np.asarray(got.reprojection_error) - z["reprojection_error"]).max() < 1e-6
  assert np.abs(np.asarray(got.view_angles) - z["view_angles"]).max() < 1e-4


@pytest.mark.parametrize("model", ["standard", "rational", "thin_prism", "tilted", "fisheye"])
def test_every_camera_model_against_opencv(model):
  scene = synthetic.make_scene(C=2, F=4, vis=0.25, seed=41, model=model)
  gt = scene["gt"]
  calib = from_scene(scene, guess=False)
  boards = [Board(p, size=(16, 22)) for p in scene["board_points"]]
  got = make_pose_table(calib.point_table, boards, calib.cameras)
  poses, ok, npts, err = pnp_oracle.make_pose_table(model, gt["K"], gt["dist"], scene["board_points"], [(16, 22, 1, 20, 3)] * scene["B"],
                                                    scene["points"], scene["valid"])
  assert ok.all() and np.array_equal(np.asarray(got.valid), ok) and np.array_equal(np.asarray(got.num_points), npts)
  assert np.abs(np.asarray(got.poses) - poses).max() < 1e-6
  assert np.abs(np.asarray(got.reprojection_error) - err).max() < 1e-6
  T_true = gt["cam_poses"][:, None, None] @ gt["frame_poses"][None, :, None] @ gt["board_poses"][None, None, :]
  assert np.abs(np.asarray(got.poses) - T_true).max() < 5e-3             # and it is the pose the scene was rendered from (0.3 px noise)


def test_april_grid_style_ids_use_the_tag_grid():
  """AprilGrid boards (aprilgrid.py:197-199): four corner ids per tag, the minimum-detections rule looks at tag ids = corner id // 4 on the
  tag grid.  Sparse views around the thresholds: the device's verdict per view must be the oracle's (and poses agree where valid)."""
  scene = synthetic.make_scene(C=2, F=6, vis=0.07, seed=43)
  gt = scene["gt"]
  calib = from_scene(scene, guess=False)
  grid = (9, 9, 4, 18, 4)                                      # 315 corner ids -> tag ids 0..78 on a 9 x 9 grid
  boards = [Board(p, size=(9, 9), min_points=18, min_rows=4, id_divisor=4) for p in scene["board_points"]]
  got = make_pose_table(calib.point_table, boards, calib.cameras)
  poses, ok, npts, err = pnp_oracle.make_pose_table("standard", gt["K"], gt["dist"], scene["board_points"], [grid] * scene["B"], scene["points"], scene["valid"])
  assert 0 < ok.sum() < ok.size                                # the scene straddles the rule
  assert np.array_equal(np.asarray(got.valid), ok) and np.array_equal(np.asarray(got.num_points), npts)
  assert np.abs(np.asarray(got.poses) - poses).max() < 1e-6
  for c, f, b in np.argwhere(ok | ~ok):                        # and the host mirror of the rule says the same
    ids = np.flatnonzero(scene["valid"][c, f, b])
    assert boards[b].has_min_detections(Table.create(ids=ids)) == bool(ok[c, f, b])


def test_minimum_detections_rule_and_bad_inputs():
  scene = synthetic.make_scene(C=2, F=3, vis=0.5, seed=42)
  calib = from_scene(scene, guess=False)
  valid = np.asarray(calib.point_table.valid).copy()
  valid[0, 0, 0, 19:] = False                                             # 19 corners < min_points
  valid[0, 1, 0] = False; valid[0, 1, 0, :32] = True                      # 32 corners, but only 2 rows of the 16-wide id grid
  valid[1, 0, 0] = False; valid[1, 0, 0, ::16] = True                     # one column
  valid[1, 1, 0] = False                                                  # nothing detected
  table = Table.create(points=calib.point_table.points, valid=valid)
  boards = [Board(p, size=(16, 22)) for p in scene["board_points"]]
  got = make_pose_table(table, boards, calib.cameras)
  want = np.ones((2, 3, 1), bool); want[0, 0] = want[0, 1] = want[1, 0] = want[1, 1] = False
  assert np.array_equal(np.asarray(got.valid), want)
  for w in zip(*np.nonzero(~want)):
    assert np.array_equal(np.asarray(got.poses)[w], np.eye(4)) and got.num_points[w] == 0 and got.reprojection_error[w] == 0
  eng = get_engine()
  det_start, det_ids, det_xy = detection_lists(table)
  bp = np.stack(scene["board_points"]); intr = np.stack([c.param_vec for c in calib.cameras])
  with pytest.raises(AssertionError): eng.pnp_views("standard", valid.shape, det_start, det_ids + 400, det_xy, bp, intr, [[16, 22, 1, 20, 3]])
  with pytest.raises(AssertionError): eng.pnp_views("standard", valid.shape, det_start[::-1].copy(), det_ids, det_xy, bp, intr, [[16, 22, 1, 20, 3]])
  with pytest.raises(NotImplementedError): eng.pnp_views("standard", valid.shape, det_start, det_ids, det_xy, bp, intr, [[80, 22, 1, 20, 3]])
