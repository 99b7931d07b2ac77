"""Wall-clock of what round 1 added without a GPU at hand: bundle adjustment under the RollingFrames / HandEye motion models and the
batched board-pose initialisation next to OpenCV's per-view calls (the oracle), on cfg2-sized synthetic scenes.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multical_b200 import synthetic
from multical_b200.board import Board
from multical_b200.calibration import from_scene
from multical_b200.motion import HandEye, RollingFrames
from multical_b200.pose_set import pose_table
from multical_b200.tables import make_pose_table
from oracle import pnp_oracle       # bench-side CPU baseline only


def timed(f, n=3):
  f(); t = time.perf_counter()
  for _ in range(n): out = f()
  return (time.perf_counter() - t) / n, out


F = int(sys.argv[1]) if len(sys.argv) > 1 else 200          # cfg2 by default
scene = synthetic.make_scene(C=4, F=F, vis=0.65, seed=0)
calib = from_scene(scene).enable(cameras=True)
t, out = timed(lambda: calib.bundle_adjust())
print(f"static   : {t * 1e3:8.2f} ms  nfev {out.last_solve.nfev} cost {out.last_solve.cost:.4f} device {out.last_solve.device_ms:.3f} ms launches {out.last_solve.kernel_launches}")
roll = calib.copy(motion=RollingFrames.init(calib.motion.pose_table))
t, out = timed(lambda: roll.bundle_adjust())
print(f"rolling  : {t * 1e3:8.2f} ms  nfev {out.last_solve.nfev} cost {out.last_solve.cost:.4f} device {out.last_solve.device_ms:.3f} ms launches {out.last_solve.kernel_launches}")
frames = np.asarray(calib.motion.poses)
he = HandEye(pose_table(frames, calib.motion.valid), np.eye(4), np.eye(4))          # arm poses = the frame estimates: G = W = identity to start
hand = calib.copy(motion=he).enable(camera_poses=False, cameras=False)
t, out = timed(lambda: hand.bundle_adjust())
print(f"hand-eye : {t * 1e3:8.2f} ms  nfev {out.last_solve.nfev} cost {out.last_solve.cost:.4f} device {out.last_solve.device_ms:.3f} ms launches {out.last_solve.kernel_launches}")

gt = from_scene(scene, guess=False)
boards = [Board(p, size=(16, 22)) for p in scene["board_points"]]
t, tab = timed(lambda: make_pose_table(gt.point_table, boards, gt.cameras))
views = int(np.asarray(tab.valid).size)
print(f"pnp gpu  : {t * 1e3:8.2f} ms for {views} views ({views / t:.0f} views/s), {int(np.asarray(tab.valid).sum())} valid")
g = scene["gt"]
t0 = time.perf_counter()
poses, ok, npts, err = pnp_oracle.make_pose_table("standard", g["K"], g["dist"], scene["board_points"], [(16, 22, 1, 20, 3)], scene["points"][:, :50], scene["valid"][:, :50])      # at most 50 frames on the host
t1 = time.perf_counter() - t0
print(f"pnp cv2  : {t1 * 1e3:8.2f} ms for {ok.size} views ({ok.size / t1:.0f} views/s, one host core)  max|dT| vs gpu {np.abs(poses - np.asarray(tab.poses)[:, :50]).max():.2e}")
