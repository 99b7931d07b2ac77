import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from multical_b200 import synthetic
from multical_b200.calibration import from_scene, get_engine
scene = synthetic.make_workload("cfg4")
for key in ("points", "valid"):
  scene[key] = torch.from_numpy(np.ascontiguousarray(scene[key])).pin_memory().numpy()
calib = from_scene(scene).enable(cameras=True)
eng = get_engine()
pts, bp = np.asarray(calib.point_table.points), calib.board_points.points
pv, vv = calib.pose_valid, np.asarray(calib.point_table.valid)
print("valid is the pinned buffer:", vv.ctypes.data == scene["valid"].ctypes.data, "points:", pts.ctypes.data == scene["points"].ctypes.data, flush=True)
for i in range(3):
  torch.cuda.synchronize(); t = time.perf_counter()
  eng.upload_dense(calib.engine_model, calib._optimize_bits(), vv, pts, bp, view_valid=pv)
  print("call %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), file=sys.stderr, flush=True)
pts32 = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).pin_memory().numpy()
for i in range(3):
  torch.cuda.synchronize(); t = time.perf_counter()
  eng.upload_dense(calib.engine_model, calib._optimize_bits(), vv, pts32, bp, view_valid=pv)
  print("f32 call %d: %.3f ms" % (i, (time.perf_counter() - t) * 1e3), file=sys.stderr, flush=True)
for name, arr in (("f64 104 MB", pts), ("f32 52 MB", pts32), ("mask 6.5 MB", vv.view(np.uint8))):
  dst = torch.empty(arr.nbytes, dtype=torch.uint8, device="cuda"); src = torch.from_numpy(arr.view(np.uint8).reshape(-1))
  for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t
  print("single H2D %-12s %.3f ms (%.1f GB/s)" % (name, dt * 1e3, arr.nbytes / dt / 1e9), file=sys.stderr, flush=True)
