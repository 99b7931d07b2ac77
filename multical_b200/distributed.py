"""Frame sharding across GPUs (one process per GPU).

Every residual depends on exactly one frame pose plus the shared parameters (motion/static_frames.py:16-25),
so frames are the natural shard: each rank holds a contiguous frame range, its corners and its frame poses, and a
replica of the shared blocks.  The only data-path exchange is the all-reduce of the reduced (shared-parameter)
normal equations inside libmcba (NCCL over NVLink); this module is the host plumbing around it."""
import numpy as np


def frame_range(F, rank, world):
  """Contiguous, balanced [start, stop) of frames owned by `rank`."""
  base, rem = divmod(F, world)
  start = rank * base + min(rank, rem)
  return start, start + base + (1 if rank < rem else 0)


def frame_owner(F, world):
  owner = np.zeros(F, np.int32)
  for r in range(world):
    a, b = frame_range(F, r, world)
    owner[a:b] = r
  return owner


def init_comm(engine, rank, world, group=None, peer_cap=None):
  """Create the engine's communicator: rank and world, then every rank exports an NVLink exchange buffer (IPC handle) and imports
  everyone else's -- the exchanges of an LM iteration run inside the solver kernel over these peer mappings (csrc/lm_kernel.cuh:
  exchange()).  `peer_cap` = doubles per slot: it must hold the reduced normal equations, n_s^2 + n_s (default 2 Mi doubles:
  n_s up to ~1400, i.e. 64 cameras with their intrinsics; 2 x world slots of that size per rank)."""
  import os
  import torch.distributed as dist
  uid = [engine.comm_unique_id() if rank == 0 else None]
  dist.broadcast_object_list(uid, src=0, group=group)
  engine.comm_init(uid[0], rank, world)
  if world > 1:
    cap = int(os.environ.get("MCBA_PEER_CAP", peer_cap or (1 << 21)))
    handles = [None] * world
    dist.all_gather_object(handles, engine.peer_export(cap), group=group)
    engine.peer_import(handles)


_dist_engines = {}       # (device, world, group id) -> Engine with a communicator and peer buffers; never the process-wide single-GPU engine


class _sharded_engine:
  """Context manager: a DEDICATED engine (own C-ABI context, communicator, NVLink exchange buffers) stands in for the process-wide
  one while a sharded call runs, and is taken out again afterwards -- a later plain `Calibration.bundle_adjust()` on this process
  must not find a context that still exchanges with the other ranks (it would wait for peers that never call)."""
  def __init__(self, rank, world, group=None, engine=None):
    self.rank, self.world, self.group, self.engine = rank, world, group, engine

  def __enter__(self):
    from . import calibration
    from .engine import Engine
    dev = calibration.default_device()
    self.dev, self.saved = dev, calibration._engines.get(dev)
    if self.engine is None:
      if getattr(self.saved, "world", 1) == self.world:
        self.engine = self.saved             # the caller installed a communicating engine itself (bench.py, tests)
      else:
        key = (dev, self.world, id(self.group))
        if key not in _dist_engines:
          eng = Engine(dev)
          init_comm(eng, self.rank, self.world, self.group)
          _dist_engines[key] = eng
        self.engine = _dist_engines[key]
    calibration._engines[dev] = self.engine
    return self.engine

  def __exit__(self, *exc):
    from . import calibration
    if self.saved is None: calibration._engines.pop(self.dev, None)
    else: calibration._engines[self.dev] = self.saved
    return False


def _slice_motion(motion, a, b):
  """Frames [a, b) of a motion model.  StaticFrames: its poses; RollingFrames (rolling_frames.py:66-150): start and end poses;
  HandEye (hand_eye.py:14-90): the fixed arm poses -- the two optimised transforms are shared parameters, replicated."""
  from .motion import MOTION_HAND_EYE, MOTION_ROLLING, motion_kind
  kind = motion_kind(motion)
  if kind == MOTION_ROLLING:
    return motion.copy(pose_start=np.asarray(motion.pose_start)[a:b], pose_end=np.asarray(motion.pose_end)[a:b],
                       valid=np.asarray(motion.valid)[a:b], names=list(motion.names)[a:b])
  if kind == MOTION_HAND_EYE:
    arm = motion.base_wrt_gripper
    return motion.copy(base_wrt_gripper=type(arm).create(poses=np.asarray(arm.poses)[a:b], valid=np.asarray(arm.valid)[a:b]), names=None)
  mt = motion.pose_table
  return motion.copy(pose_table=type(mt).create(poses=np.asarray(mt.poses)[a:b], valid=np.asarray(mt.valid)[a:b]), names=None)


def shard_calibration(calib, rank, world):
  """Local view of a Calibration: this rank's frames only (point table, inlier mask, the motion model's per-frame state)."""
  F = calib.size.rig_poses
  a, b = frame_range(F, rank, world)
  pt = calib.point_table
  make = getattr(type(pt), "create")
  local_pt = make(points=np.asarray(pt.points)[:, a:b], valid=np.asarray(pt.valid)[:, a:b])
  mask = None if calib.inlier_mask is None else calib.inlier_mask[:, a:b]
  return calib.copy(point_table=local_pt, motion=_slice_motion(calib.motion, a, b), inlier_mask=mask), (a, b)


def merge_motion(full_motion, local_motion, F, rank, world, group=None, comm=None):
  """The full motion model after a sharded solve: per-frame state all-gathered in frame order, shared state from the local result
  (every rank solves the shared system redundantly on bit-identical data)."""
  from .motion import MOTION_HAND_EYE, MOTION_ROLLING, motion_kind
  kind = motion_kind(full_motion)
  if kind == MOTION_ROLLING:
    return full_motion.copy(pose_start=gather_frames(local_motion.pose_start, F, rank, world, group, comm),
                            pose_end=gather_frames(local_motion.pose_end, F, rank, world, group, comm))
  if kind == MOTION_HAND_EYE:
    return full_motion.copy(world_wrt_base=local_motion.world_wrt_base, gripper_wrt_camera=local_motion.gripper_wrt_camera)
  mt = full_motion.pose_table
  return full_motion.copy(pose_table=type(mt).create(poses=gather_frames(local_motion.poses, F, rank, world, group, comm), valid=np.asarray(mt.valid)))


def gather_frames(local_frame_poses, F, rank, world, group=None, comm=None):
  """All-gather the per-rank frame poses back into the full [F,4,4] table (host side, after the solve); through `comm.all_gather`
  when a communicator object is given (TorchComm below, or the thread communicator of the CPU tests), else torch.distributed."""
  if comm is not None:
    parts = comm.all_gather(np.asarray(local_frame_poses))
  else:
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, np.asarray(local_frame_poses), group=group)
  out = np.concatenate(parts, axis=0)
  assert out.shape[0] == F
  return out


def bundle_adjust(calib, group=None, **kwargs):
  """Multi-GPU `Calibration.bundle_adjust`: call on every rank with the same full Calibration; returns the same
  full, updated Calibration on every rank."""
  import torch.distributed as dist
  rank, world = dist.get_rank(group), dist.get_world_size(group)
  local, (a, b) = shard_calibration(calib, rank, world)
  with _sharded_engine(rank, world, group):
    out_local = local.bundle_adjust(**kwargs)
  motion = merge_motion(calib.motion, out_local.motion, calib.size.rig_poses, rank, world, group)
  out = calib.copy(cameras=out_local.cameras, camera_poses=out_local.camera_poses, board_poses=out_local.board_poses, motion=motion)
  out.__dict__["last_solve"] = out_local.last_solve
  return out


class TorchComm:
  """The two host-side collectives of the sharded outlier loop over torch.distributed (small Python objects: a few hundred order
  statistics per round; the bundle adjustments themselves exchange over NCCL / NVLink inside libmcba)."""
  def __init__(self, group=None):
    import torch.distributed as dist
    self.dist, self.group = dist, group
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

  def all_gather(self, obj):
    out = [None] * self.world
    self.dist.all_gather_object(out, obj, group=self.group)
    return out

  def all_reduce_sum(self, arr):
    return np.sum(self.all_gather(np.asarray(arr)), axis=0)


def adjust_outliers(calib, comm=None, group=None, num_adjustments=3, select_scale=None, select_outliers=None, **kwargs):
  """Multi-GPU `Calibration.adjust_outliers` (calibration.py:254-268) with every rank's frame shard of the point table resident on
  its GPU: call on every rank with the same full Calibration; returns the same full, updated Calibration (poses, intrinsics,
  inlier mask) on every rank.  Selectors must be quantile rules (`select_threshold`) or None."""
  from .calibration import get_engine
  comm = comm or TorchComm(group)
  rank, world = comm.rank, comm.world
  local, (a, b) = shard_calibration(calib, rank, world)
  eng = get_engine()
  if getattr(eng, "world", 1) == world:        # the caller's engine already talks to the other ranks (thread ranks of the CPU tests, bench.py)
    out_local = local._adjust_outliers_resident(num_adjustments, select_scale, select_outliers, comm=comm, **kwargs)
  else:
    with _sharded_engine(rank, world, group):
      out_local = local._adjust_outliers_resident(num_adjustments, select_scale, select_outliers, comm=comm, **kwargs)
  F = calib.size.rig_poses
  motion = merge_motion(calib.motion, out_local.motion, F, rank, world, group, comm)
  mask = None
  if out_local.inlier_mask is not None:
    mask = np.concatenate(comm.all_gather(np.asarray(out_local.inlier_mask)), axis=1)          # frames are axis 1 of [C,F,B,P]
  out = calib.copy(cameras=out_local.cameras, camera_poses=out_local.camera_poses, board_poses=out_local.board_poses, motion=motion,
                   inlier_mask=mask)
  out.__dict__["last_solve"] = out_local.last_solve
  return out
