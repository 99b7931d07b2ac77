"""Frame sharding: host logic on CPU with world_size-2 gloo; the NCCL path itself on >=2 GPUs (-m gpu)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from multical_b200 import distributed as mdist
from multical_b200 import synthetic
from multical_b200.calibration import from_scene


def test_frame_ranges_partition_all_frames():
  for F in (1, 7, 200, 1001):
    for world in (1, 2, 3, 8):
      ranges = [mdist.frame_range(F, r, world) for r in range(world)]
      assert ranges[0][0] == 0 and ranges[-1][1] == F
      assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
      sizes = [b - a for a, b in ranges]
      assert max(sizes) - min(sizes) <= 1
      owner = mdist.frame_owner(F, world)
      assert all((owner[a:b] == r).all() for r, (a, b) in enumerate(ranges))


def test_shards_cover_every_corner_exactly_once():
  scene = synthetic.make_scene(C=3, F=7, vis=0.5, seed=5)
  calib = from_scene(scene).enable(cameras=True)
  total, frames = 0, []
  for r in range(3):
    local, (a, b) = mdist.shard_calibration(calib, r, 3)
    assert local.size.rig_poses == b - a and local.size.cameras == 3
    assert np.array_equal(local.point_table.points, scene["points"][:, a:b])
    assert np.array_equal(local.motion.poses, calib.motion.poses[a:b])
    # shared blocks are replicated, frame block is local
    assert local.param_vec.size == calib.param_vec.size - 6 * (7 - (b - a))
    total += int(local.inliers.sum()); frames += list(range(a, b))
  assert total == int(calib.inliers.sum()) and frames == list(range(7))


def test_motion_models_shard_by_frame():
  """RollingFrames: both pose sets are per-frame state (12 parameters per local frame); HandEye: only the arm poses are sliced, the 12
  optimised parameters are shared and replicated."""
  from multical_b200.motion import HandEye, RollingFrames
  from multical_b200.pose_set import pose_table
  scene = synthetic.make_scene(C=2, F=7, vis=0.5, seed=6)
  calib = from_scene(scene)
  end = calib.motion.poses.copy(); end[:, 0, 3] += 0.01
  roll = calib.copy(motion=RollingFrames(calib.motion.poses, end, calib.motion.valid, [str(i) for i in range(7)]))
  hand = calib.copy(motion=HandEye(pose_table(calib.motion.poses, calib.motion.valid), np.eye(4), np.eye(4)))
  for r in range(3):
    a, b = mdist.frame_range(7, r, 3)
    lr, _ = mdist.shard_calibration(roll, r, 3)
    assert np.array_equal(lr.motion.pose_start, roll.motion.pose_start[a:b]) and np.array_equal(lr.motion.pose_end, end[a:b])
    assert lr.param_vec.size == roll.param_vec.size - 12 * (7 - (b - a)) and lr.motion.names == [str(i) for i in range(a, b)]
    lh, _ = mdist.shard_calibration(hand, r, 3)
    assert np.array_equal(lh.motion.base_wrt_gripper.poses, calib.motion.poses[a:b]) and lh.motion.size == b - a
    assert lh.param_vec.size == hand.param_vec.size                      # nothing per-frame in the hand-eye parameter vector
    assert np.array_equal(lh.motion.poses, hand.motion.poses[a:b])


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["MCBA_ROOT"])
from multical_b200 import distributed as mdist, synthetic
from multical_b200.calibration import from_scene
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scene = synthetic.make_scene(C=2, F=5, vis=0.5, seed=9)
calib = from_scene(scene)
local, (a, b) = mdist.shard_calibration(calib, rank, world)
# each rank perturbs its own frames; the gather must reassemble them in frame order on every rank
mine = np.asarray(local.motion.poses).copy(); mine[:, 0, 3] += 100.0 * (rank + 1)
full = mdist.gather_frames(mine, 5, rank, world)
expect = np.asarray(calib.motion.poses).copy()
for r in range(world):
  lo, hi = mdist.frame_range(5, r, world); expect[lo:hi, 0, 3] += 100.0 * (r + 1)
assert np.array_equal(full, expect)
# a rolling-shutter model: start and end poses are gathered separately, both in frame order
from multical_b200.motion import HandEye, RollingFrames
from multical_b200.pose_set import pose_table
end = np.asarray(calib.motion.poses).copy(); end[:, 1, 3] += 0.5
roll = calib.copy(motion=RollingFrames(calib.motion.poses, end, calib.motion.valid, [str(i) for i in range(5)]))
lroll, _ = mdist.shard_calibration(roll, rank, world)
moved = lroll.motion.copy(pose_start=mine, pose_end=np.asarray(lroll.motion.pose_end) + (rank + 1))
merged = mdist.merge_motion(roll.motion, moved, 5, rank, world)
expect_end = end.copy()
for r in range(world):
  lo, hi = mdist.frame_range(5, r, world); expect_end[lo:hi] += r + 1
assert np.array_equal(merged.pose_start, expect) and np.array_equal(merged.pose_end, expect_end)
# hand-eye: the optimised pair is shared state, the arm table stays whole
hand = calib.copy(motion=HandEye(pose_table(calib.motion.poses, calib.motion.valid), np.eye(4), np.eye(4)))
lhand, _ = mdist.shard_calibration(hand, rank, world)
W = np.eye(4); W[0, 3] = 0.25
mh = mdist.merge_motion(hand.motion, lhand.motion.copy(world_wrt_base=W), 5, rank, world)
assert np.array_equal(mh.world_wrt_base, W) and mh.size == 5
# the unique-id broadcast used for the NCCL communicator (object broadcast from rank 0)
uid = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
assert uid[0] == bytes(range(128))
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_gather_and_broadcast_with_gloo_world2(tmp_path):
  script = tmp_path / "worker.py"; script.write_text(_GLOO_WORKER)
  env = dict(os.environ, MCBA_ROOT=ROOT)
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()), str(script)]
  out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
  assert out.returncode == 0, out.stdout + out.stderr
  assert out.stdout.count("ok") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", ["0", "1"])
def test_two_gpu_solve_matches_single_gpu(fuse):
  """fuse = "1": MCBA_FUSE=1 -- five exchanges per LM iteration, four of them as kernel tails (csrc/peer_allreduce.cuh peer_allreduce_block)."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
  env = dict(os.environ); env.pop("MCBA_FUSE", None)
  if fuse == "1": env["MCBA_FUSE"] = "1"
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
  assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
  assert "MULTI_GPU_OK" in out.stdout
