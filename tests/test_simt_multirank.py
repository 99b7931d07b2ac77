"""CPU (-m "not gpu"): the multi-GPU path of mcba_solve on the SIMT interpreter -- two ranks as two host threads of this process, each
with its own C-ABI context, frames sharded between them (multical_b200/distributed.py), the exchanges of every LM iteration running
inside the real k_lm kernel (csrc/lm_kernel.cuh: exchange()) over "peer" buffers -- the IPC handle of the interpreter build carries the pointer.  The sharded solve must reproduce the single-rank solve, the same
assertion scripts/multi_gpu_check.py makes on real GPUs.  What this cannot show: NVLink ordering / visibility -- that is hardware."""
import threading

import numpy as np
import pytest

import test_gpu_motion as gm
import test_gpu_parity as gp
from multical_b200 import _native, calibration
from multical_b200 import distributed as mdist
from multical_b200.engine import Engine


@pytest.fixture(scope="module")
def simt_library():
  import simt
  return simt.build()


@pytest.fixture(autouse=True)
def on_the_interpreter(simt_library, monkeypatch):
  monkeypatch.setattr(_native, "LIB_PATH", simt_library)
  monkeypatch.setattr(_native, "_lib", None)
  monkeypatch.setattr(_native, "_allow_interpreter", True)
  monkeypatch.setattr(calibration, "_engines", {})
  yield
  for eng in calibration._engines.values(): eng.close()


def sharded_solve(calib, world, peer, monkeypatch, **kwargs):
  """Every rank = one thread with its own engine; returns the per-rank results (Calibration of the shard, SolveInfo)."""
  tls = threading.local()
  engines = [Engine(0) for _ in range(world)]
  uid = engines[0].comm_unique_id()                       # also loads the (stand-in) NCCL entry points before the threads start
  handles, results, errors = [None] * world, [None] * world, []
  gate = threading.Barrier(world)

  def rank_main(rank):
    try:
      tls.engine = eng = engines[rank]
      eng.comm_init(uid, rank, world)
      if peer:
        handles[rank] = eng.peer_export(1 << 14)
        gate.wait()
        eng.peer_import(handles)
      local, (a, b) = mdist.shard_calibration(calib, rank, world)
      out = local.bundle_adjust(**kwargs)
      results[rank] = (out, out.last_solve, (a, b))
    except BaseException as e:                            # a failing rank must not leave the others waiting in a rendezvous forever
      errors.append(e)
      gate.abort()
      raise

  threads = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
  with monkeypatch.context() as m:
    m.setattr(calibration, "get_engine", lambda device=None: tls.engine)      # one engine per rank instead of one per device
    for t in threads: t.start()
    for t in threads: t.join(timeout=600)
  assert not errors, errors
  assert all(not t.is_alive() for t in threads), "a rank is stuck in an exchange"
  for eng in engines: eng.close()
  return results


def check_against_single(calib, results):
  single = calib.bundle_adjust()
  ref = single.last_solve
  for out, res, (a, b) in results:
    assert res.nfev == ref.nfev and res.status == ref.status
    assert abs(res.cost - ref.cost) <= 1e-9 * ref.cost
    assert np.abs(np.asarray(out.camera_poses.poses) - np.asarray(single.camera_poses.poses)).max() < 1e-6      # shared blocks: every rank has them (ftol 1e-4 solves)
    assert np.abs(np.stack([c.param_vec for c in out.cameras]) - np.stack([c.param_vec for c in single.cameras])).max() < 1e-6
  r0, r1 = results[0][1], results[1][1]
  assert r0.cost == r1.cost and np.array_equal(np.array(r0.log, float), np.array(r1.log, float), equal_nan=True)       # bit-identical scalar logic on every rank
  return single


@pytest.mark.parametrize("peer", [True])
@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6"])
def test_two_ranks_reproduce_the_single_rank_solve(name, peer, monkeypatch):
  scene, z, calib, prob = gp.make(name)
  results = sharded_solve(calib, 2, peer, monkeypatch)
  single = check_against_single(calib, results)
  frames = np.concatenate([np.asarray(out.motion.poses) for out, _, _ in results])
  assert np.abs(frames - np.asarray(single.motion.poses)).max() < 1e-6


def test_four_ranks_with_uneven_shards(monkeypatch):
  """Six frames over four ranks (2, 2, 1, 1): rank-ordered reductions with more than one peer."""
  scene, z, calib, prob = gp.make("cube3_3x6")
  single = calib.bundle_adjust().last_solve
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  results = sharded_solve(calib, 4, True, monkeypatch)
  assert [b - a for _, _, (a, b) in results] == [2, 2, 1, 1]
  for out, res, _ in results:
    assert res.nfev == single.nfev and abs(res.cost - single.cost) <= 1e-9 * single.cost
    assert res.cost == results[0][1].cost


@pytest.mark.parametrize("name", gm.CASES)
def test_two_ranks_under_the_motion_models(name, monkeypatch):
  z, calib, prob = gm.make(name)
  results = sharded_solve(calib, 2, True, monkeypatch, max_iterations=8)
  single = calib.bundle_adjust(max_iterations=8)
  for out, res, _ in results:
    assert res.nfev == single.last_solve.nfev
    assert abs(res.cost - single.last_solve.cost) <= 1e-8 * single.last_solve.cost


# ---- the outlier loop with the point table sharded over the ranks (multical_b200/distributed.py adjust_outliers) ------------------
class ThreadComm:
  """all_gather / all_reduce_sum between the rank threads of one process (what TorchComm does over torch.distributed)."""
  def __init__(self, world):
    self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)
    self.tls = threading.local()

  rank = property(lambda self: self.tls.rank)

  def all_gather(self, obj):
    self.slots[self.tls.rank] = obj
    self.barrier.wait()
    out = list(self.slots)
    self.barrier.wait()
    return out

  def all_reduce_sum(self, arr):
    return np.sum(self.all_gather(np.asarray(arr)), axis=0)


def test_merged_order_statistics_are_numpy_on_the_union():
  """Host logic of the distributed quantile (outliers.merged_order_statistics) with numpy arrays standing in for the device-side sorted
  errors: exact order statistics of the union for uneven shards, ties, an empty shard and ranks at both ends."""
  from multical_b200.outliers import merged_order_statistics, quantile_from_sorted
  rng = np.random.default_rng(0)
  for world, sizes in [(2, [1000, 37]), (3, [500, 0, 1200]), (4, [64, 64, 64, 64])]:
    shards = [np.sort(np.round(rng.gamma(2.0, 0.3, n), 2)) for n in sizes]              # rounding makes ties
    union = np.sort(np.concatenate(shards)); N = union.size
    ranks = np.array([0, 1, N // 3, N // 2, N - 2, N - 1])
    comm = ThreadComm(world); results = [None] * world
    def main(r):
      comm.tls.rank = r
      got = merged_order_statistics(comm, lambda lr: shards[r][lr], lambda v: np.searchsorted(shards[r], v, side="left"), sizes[r], ranks, splitters=16)
      q = quantile_from_sorted(lambda rk: merged_order_statistics(comm, lambda lr: shards[r][lr], lambda v: np.searchsorted(shards[r], v, side="left"),
                                                                  sizes[r], rk, splitters=16), N, np.array([0.0, 0.25, 0.5, 0.75, 1.0]))
      results[r] = (got, q)
    threads = [threading.Thread(target=main, args=(r,), daemon=True) for r in range(world)]
    for t in threads: t.start()
    for t in threads: t.join(timeout=60)
    for got, q in results:
      assert np.array_equal(got, union[ranks])
      assert np.array_equal(q, np.quantile(union, [0.0, 0.25, 0.5, 0.75, 1.0]))


def test_sharded_outlier_loop_equals_the_single_rank_loop(monkeypatch):
  from multical_b200 import synthetic
  from multical_b200.calibration import from_scene, select_threshold
  scene = synthetic.make_scene(C=3, F=8, vis=0.5, seed=61, outlier_fraction=0.02)
  calib = from_scene(scene).enable(cameras=True)
  kw = dict(num_adjustments=2, select_outliers=select_threshold(quantile=0.75, factor=4), select_scale=select_threshold(quantile=0.5, factor=3), loss="soft_l1")
  single = calib.adjust_outliers(**kw)
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()

  world = 2
  comm = ThreadComm(world)
  tls = threading.local()
  engines = [Engine(0) for _ in range(world)]
  uid = engines[0].comm_unique_id()
  handles, results, errors = [None] * world, [None] * world, []
  gate = threading.Barrier(world)
  def rank_main(rank):
    try:
      comm.tls.rank = rank
      tls.engine = eng = engines[rank]
      eng.comm_init(uid, rank, world)
      handles[rank] = eng.peer_export(1 << 14); gate.wait(); eng.peer_import(handles)
      results[rank] = mdist.adjust_outliers(calib, comm=comm, **kw)
    except BaseException as e:
      errors.append(e); gate.abort(); comm.barrier.abort(); raise
  threads = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
  with monkeypatch.context() as m:
    m.setattr(calibration, "get_engine", lambda device=None: tls.engine)
    for t in threads: t.start()
    for t in threads: t.join(timeout=600)
  assert not errors, errors
  for eng in engines: eng.close()
  for out in results:
    assert np.array_equal(out.inlier_mask, single.inlier_mask) and out.inlier_mask.sum() < calib.valid.sum()
    assert abs(out.last_solve.cost - single.last_solve.cost) <= 1e-8 * single.last_solve.cost
    assert np.abs(np.asarray(out.motion.poses) - np.asarray(single.motion.poses)).max() < 1e-6
    assert np.abs(np.asarray(out.camera_poses.poses) - np.asarray(single.camera_poses.poses)).max() < 1e-6
