"""Dynamic synchronisation counts of one small solve on the SIMT interpreter (tests/simt): launches, block-wide barriers completed and
warp collectives per kernel -- the serialisation points of the latency-bound small kernels, independent of any clock.  Run twice:
  SIMT_STATS=out_default.tsv python scripts/simt_sync_counts.py
  SIMT_STATS=out_candidates.tsv MCBA_CHOL=blocked MCBA_FUSE=1 python scripts/simt_sync_counts.py
Test infrastructure only (needs g++, no GPU); 4 cameras so that the reduced system has cfg2's size (n_s = 70)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simt
from multical_b200 import _native, synthetic
_native.LIB_PATH = simt.build(); _native._allow_interpreter = True
from multical_b200.calibration import from_scene
scene = synthetic.make_scene(C=4, F=8, vis=0.3, seed=0)
calib = from_scene(scene).enable(cameras=True)
r = calib.bundle_adjust().last_solve
print("n_s = 70, nfev", r.nfev, "launches", r.kernel_launches, "cost %.6f" % r.cost)
