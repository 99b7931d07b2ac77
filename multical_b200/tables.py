"""`make_pose_table` with the reference's signature and result (multical/tables.py:44-66) -- the per-view board poses that
initialise the bundle adjustment (workspace.py:196-226) -- computed in one launch on the GPU (include/mcba.h mcba_pnp_views)
instead of C*F*B calls of `board.estimate_pose_points` (board/common.py:36-47).  The pose-graph step that follows
(`tables.initialise_poses`, tables.py:354-377) is small host work and stays with the reference."""
import numpy as np
from scipy.spatial.transform import Rotation

from .board import stack_boards
from .calibration import get_engine
from .camera import engine_model_of
from .structs import Table


def board_grid_of(board):
  """{id-grid width, height, id divisor, min_points, min_rows} of `board.has_min_detections` (charuco.py:104-106: corner ids on
  the (w, h) = board.size grid; aprilgrid.py:197-199: tag ids = corner ids // 4)."""
  w, h = (int(v) for v in board.size)
  return [w, h, int(getattr(board, "id_divisor", 1)), int(board.min_points), int(board.min_rows)]


def detection_lists(point_table):
  """Dense [C,F,B,P] table -> CSR detection lists in (c, f, b) order: what tables.sparse_points (tables.py:34-36) extracts view by view."""
  valid = np.asarray(point_table.valid)
  pts = np.asarray(point_table.points)
  counts = valid.reshape(-1, valid.shape[-1]).sum(axis=1)
  det_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
  idx = np.argwhere(valid)
  return det_start, idx[:, 3].astype(np.int32), np.ascontiguousarray(pts[valid])


def make_pose_table(point_table, boards, cameras, exclude_bad_poses=False, pose_error_limit=1.0):
  """tables.py:44-66.  Returns Table(poses [C,F,B,4,4], valid, num_points, reprojection_error, view_angles [C,F,B,3]); a view without
  the minimum detections, or (exclude_bad_poses) with error above pose_error_limit, is the reference's invalid_pose."""
  valid = np.asarray(point_table.valid)
  Cn, F, B, P = valid.shape
  models = {engine_model_of(c) for c in cameras}
  assert len(models) == 1, f"all cameras must share one model, got {models}"
  det_start, det_ids, det_xy = detection_lists(point_table)
  bp, _ = stack_boards(boards)
  intr = np.stack([np.asarray(c.param_vec, np.float64) for c in cameras])
  grid = np.array([board_grid_of(b) for b in boards], np.int32)
  poses, err, npts, ok = get_engine().pnp_views(models.pop(), (Cn, F, B, P), det_start, det_ids, det_xy, bp, intr, grid)
  if exclude_bad_poses:
    ok = ok & ~(err > pose_error_limit)
  poses = np.where(ok[..., None, None], poses, np.eye(4))
  angles = np.zeros((Cn, F, B, 3))
  if ok.any():                                       # rtvec.rtvec_to_euler (transform/rtvec.py:55-58)
    angles[ok] = Rotation.from_matrix(poses[ok][:, :3, :3]).as_euler("xyz", degrees=True)
  return Table.create(poses=poses, valid=ok, num_points=np.where(ok, npts, 0), reprojection_error=np.where(ok, err, 0.0), view_angles=angles)
