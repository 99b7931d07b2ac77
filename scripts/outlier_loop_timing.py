"""Wall-clock of Calibration.adjust_outliers: host loop (dense table re-uploaded around every step, errors to the host)
against the resident point table (developer diagnostics, GPU box).
usage: outlier_loop_timing.py [workload] [repeats] [both|resident]   (resident: only that loop, e.g. under ncu)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multical_b200 import synthetic
from multical_b200.calibration import from_scene, select_threshold

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "both"
args = dict(synthetic.WORKLOADS[wl]); args.update(seed=7, outlier_fraction=0.02)
scene = synthetic.make_scene(**args)
calib = from_scene(scene).enable(cameras=True)
kw = dict(num_adjustments=3, select_outliers=select_threshold(quantile=0.75, factor=5.0),
          select_scale=select_threshold(quantile=0.75, factor=2.0), loss="soft_l1")


def run(host):
  if host: os.environ["MCBA_HOST_OUTLIERS"] = "1"
  else: os.environ.pop("MCBA_HOST_OUTLIERS", None)
  calib.adjust_outliers(**kw)                                           # warm-up (allocations)
  times = []
  for _ in range(repeats):
    t = time.perf_counter()
    out = calib.adjust_outliers(**kw)
    times.append(time.perf_counter() - t)
  return out, float(np.median(times)) * 1e3


res, res_ms = run(False)
if which == "resident":
  print(json.dumps(dict(workload=wl, corners=int(calib.valid.sum()), resident_loop_ms=round(res_ms, 3), inliers=int(res.inlier_mask.sum()))))
  sys.exit(0)
host, host_ms = run(True)
print(json.dumps(dict(workload=wl, corners=int(calib.valid.sum()), table_bytes=int(calib.valid.size * 17), adjustments=3,
                      host_loop_ms=round(host_ms, 3), resident_loop_ms=round(res_ms, 3), speedup=round(host_ms / res_ms, 2),
                      same_inliers=bool(np.array_equal(host.inlier_mask, res.inlier_mask)),
                      inliers=int(res.inlier_mask.sum()),
                      cost_rel_diff=abs(host.last_solve.cost - res.last_solve.cost) / host.last_solve.cost)))
