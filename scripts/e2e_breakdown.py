"""Where does the end-to-end time of Calibration.bundle_adjust() go? (developer diagnostics, GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multical_b200 import synthetic
from multical_b200.calibration import from_scene, get_engine
from multical_b200.engine import format_log

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scene = synthetic.make_workload(wl)
for key in ("points", "valid"):
  scene[key] = torch.from_numpy(np.ascontiguousarray(scene[key])).pin_memory().numpy()
eng = get_engine()
def T(label, f, n=20):
  f(); torch.cuda.synchronize()
  t = time.perf_counter()
  for _ in range(n): out = f()
  torch.cuda.synchronize()
  print(f"{label:40s} {(time.perf_counter() - t) / n * 1e3:8.3f} ms", flush=True)
  return out
calib = from_scene(scene).enable(cameras=True)
T("from_scene + enable", lambda: from_scene(scene).enable(cameras=True))
T("inliers (valid mask)", lambda: from_scene(scene).inliers)
T("board_points", lambda: from_scene(scene).board_points)
T("_optimize_bits + engine_model", lambda: (from_scene(scene).enable(cameras=True)._optimize_bits(), from_scene(scene).engine_model))
m, pts, bp = calib.inliers, np.asarray(calib.point_table.points), calib.board_points.points
T("upload_dense (inliers built on the host)", lambda: eng.upload_dense(calib.engine_model, calib._optimize_bits(), m, pts, bp))
pv, vv = calib.pose_valid, np.asarray(calib.point_table.valid)
T("upload_dense (valid + view_valid, f64)", lambda: eng.upload_dense(calib.engine_model, calib._optimize_bits(), vv, pts, bp, view_valid=pv))
pts32 = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).pin_memory().numpy()
T("upload_dense (valid + view_valid, f32 table)", lambda: eng.upload_dense(calib.engine_model, calib._optimize_bits(), vv, pts32, bp, view_valid=pv))
t_ = torch.empty(pts.nbytes, dtype=torch.uint8, device="cuda"); src_ = torch.from_numpy(pts.view(np.uint8).reshape(-1))
T("plain H2D of the f64 table (torch, pinned)", lambda: t_.copy_(src_, non_blocking=True))
eng.upload_dense(calib.engine_model, calib._optimize_bits(), m, pts, bp)
mats = np.concatenate([calib.camera_poses.poses, calib.board_poses.poses, calib.motion.poses])
intr = np.stack([c.param_vec for c in calib.cameras])
T("set_state_matrices", lambda: eng.set_state_matrices(mats, intr))
def solve():
  eng.set_state_matrices(mats, intr); return eng.solve(ftol=1e-4, max_nfev=100)
res = T("set_state_matrices + solve", solve)
T("get_state_matrices", lambda: eng.get_state_matrices())
T("_with_engine_state", lambda: calib._with_engine_state(eng))
print("   device_ms", res.device_ms, "launches", res.kernel_launches)
T("format_log + info", lambda: [l for l in format_log(res.log)])
T("eng.param_vec", lambda: eng.param_vec)
x = eng.param_vec
T("with_param_vec", lambda: calib.with_param_vec(x))
T("full bundle_adjust (fresh object)", lambda: from_scene(scene).enable(cameras=True).bundle_adjust())
