"""CPU (-m "not gpu"): the multi-GPU path of mcba_solve on the SIMT interpreter -- two ranks as two host threads of this process, each
with its own C-ABI context, frames sharded between them (multical_b200/distributed.py), the exchanges of every LM iteration going
through the real k_peer_allreduce kernel over "peer" buffers (the IPC handle of the interpreter build carries the pointer) or through
an in-process stand-in for NCCL (tests/simt/shim/cuda_runtime.h).  The sharded solve must reproduce the single-rank solve, the same
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
    assert np.abs(np.asarray(out.camera_poses.poses) - np.asarray(single.camera_poses.poses)).max() < 1e-7      # shared blocks: every rank has them
    assert np.abs(np.stack([c.param_vec for c in out.cameras]) - np.stack([c.param_vec for c in single.cameras])).max() < 1e-6
  r0, r1 = results[0][1], results[1][1]
  assert r0.cost == r1.cost and np.array_equal(np.array(r0.log, float), np.array(r1.log, float), equal_nan=True)       # bit-identical scalar logic on every rank
  return single


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6"])
def test_two_ranks_reproduce_the_single_rank_solve(name, peer, monkeypatch):
  scene, z, calib, prob = gp.make(name)
  results = sharded_solve(calib, 2, peer, monkeypatch)
  single = check_against_single(calib, results)
  frames = np.concatenate([np.asarray(out.motion.poses) for out, _, _ in results])
  assert np.abs(frames - np.asarray(single.motion.poses)).max() < 1e-7


@pytest.mark.parametrize("name", gm.CASES)
def test_two_ranks_under_the_motion_models(name, monkeypatch):
  z, calib, prob = gm.make(name)
  results = sharded_solve(calib, 2, True, monkeypatch, max_iterations=8)
  single = calib.bundle_adjust(max_iterations=8)
  for out, res, _ in results:
    assert res.nfev == single.last_solve.nfev
    assert abs(res.cost - single.last_solve.cost) <= 1e-8 * single.last_solve.cost


@pytest.mark.parametrize("peer", [True, False])
def test_two_ranks_with_the_merged_first_exchange(peer, monkeypatch):
  """MCBA_FUSE=1 on several ranks: the frame parts of the scaling sums travel with g_s / diag(H_ss) / cost (k_scale_part), one
  exchange per LM iteration fewer; the solve must still be the single-rank solve."""
  scene, z, calib, prob = gp.make("cube3_3x6")
  single = calib.bundle_adjust().last_solve                 # default kernels, one rank
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  monkeypatch.setenv("MCBA_FUSE", "1")
  results = sharded_solve(calib, 2, peer, monkeypatch)
  for out, res, _ in results:
    assert res.nfev == single.nfev and res.status == single.status
    assert abs(res.cost - single.cost) <= 1e-9 * single.cost
  assert results[0][1].cost == results[1][1].cost
  assert results[0][1].kernel_launches < sharded_solve(calib, 2, peer, monkeypatch_env_off(monkeypatch))[0][1].kernel_launches


def monkeypatch_env_off(monkeypatch):
  monkeypatch.delenv("MCBA_FUSE")
  return monkeypatch
