#!/usr/bin/env python
"""bench.py — reprojection residuals/sec and LM iterations/sec of the bundle-adjustment hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

One "step" = one full `bundle_adjust()` (calibration.py:199-212 semantics: ftol=1e-4, max_nfev=100, linear
loss) over one synthetic scene of the BASELINE.json configuration `--workload` (default cfg2 = configs[1],
4 cameras x 200 frames x charuco_16x22, ~160k corners).  Metric (both arms, same definition):
    residuals/s = N_corners * (nfev + njev) / time      1 residual = one inlier corner (2 scalars),
    nfev/njev = cost and Jacobian evaluations as the solver reports them (scipy's res.nfev/res.njev for the
    reference arm; its finite-difference sub-evaluations are NOT counted, they are an artefact of its Jacobian).
`value`  : solves timed on the device with the packed problem already resident in HBM (CUDA events).
`e2e`    : the same through the public API `Calibration.bundle_adjust()` from host numpy buffers: packing,
           H2D, solve, D2H of the parameter vector all inside the timed region (wall clock, device synced).
`--impl reference` times the reference's CPU algorithm (oracle/ba_oracle.py: dense numpy evaluate + the identical
scipy.optimize.least_squares call) on a bounded frame-subsample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "reprojection_residuals_per_sec", "residuals/s"
BA_KW = dict(tolerance=1e-4, max_iterations=100, loss="linear", f_scale=1.0)


def subsample_frames(scene, frames):
  s = dict(scene)
  s["F"] = len(frames)
  s["points"] = scene["points"][:, frames]; s["valid"] = scene["valid"][:, frames]
  s["frame_valid"] = scene["frame_valid"][frames]
  for k in ("init", "gt"):
    d = dict(scene[k]); d["frame_poses"] = scene[k]["frame_poses"][frames]; s[k] = d
  return s


def cpu_reference_step(scene):
  """One bundle_adjust of the reference algorithm (oracle port) on the host cores."""
  from oracle.ba_oracle import Problem
  prob = Problem.from_scene(scene, optimize=dict(cameras=True))
  t = time.perf_counter()
  _, res = prob.bundle_adjust(tolerance=BA_KW["tolerance"], max_iterations=BA_KW["max_iterations"], loss=BA_KW["loss"])
  dt = time.perf_counter() - t
  n = int(prob.inliers.sum())
  return n * (res.nfev + res.njev) / dt, dt, res, n


class ClockSampler(threading.Thread):
  """SM clock and throttle reasons sampled DURING the timed regions: NVML (1 ms period) when importable, else nvidia-smi."""
  REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

  def __init__(self, index=0):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.source = index, [], set(), False, None, None

  def _run_nvml(self):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
    self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
    get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
    self.source = "nvml"
    while not self.stop_flag:
      self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
      mask = int(get_reasons(h))
      for bit, name in self.REASONS.items():
        if mask & bit: self.reasons.add(name)
      time.sleep(0.001)

  def _run_smi(self):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    self.source = "nvidia-smi"
    while not self.stop_flag:
      try:
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0])); self.max_mhz = float(out[1])
        for n, v in zip(names, out[2:]):
          if "Active" in v and "Not" not in v: self.reasons.add(n)
      except Exception:
        pass
      time.sleep(0.05)

  def run(self):
    try:
      self._run_nvml()
    except Exception:
      if not self.stop_flag: self._run_smi()

  def summary(self):
    return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                reasons=sorted(self.reasons), samples=len(self.samples), source=self.source)


def measured_peak():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
  p = os.path.join(ROOT, "profiles", "traffic.json")
  if os.path.exists(p):
    return json.load(open(p)).get(workload)
  return None


def run_reference(args):
  from multical_b200 import synthetic
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  scene = synthetic.make_workload(args.workload, seed=args.seed)
  nf = min(scene["F"], args.ref_frames)
  sample = subsample_frames(scene, np.arange(nf))
  for _ in range(args.warmup):
    cpu_reference_step(subsample_frames(scene, np.arange(min(4, nf))))
  tot_res, tot_t, nfev, njev = 0.0, 0.0, 0, 0
  for _ in range(args.steps):
    v, dt, res, n = cpu_reference_step(sample)
    tot_res += n * (res.nfev + res.njev); tot_t += dt; nfev += res.nfev; njev += res.njev
  value = tot_res / tot_t
  line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=1e3 * tot_t / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
              data="synthetic", impl="reference",
              config=dict(workload=args.workload, sample=f"first {nf} of {scene['F']} frames, full bundle_adjust per step", corners=n),
              lm_iters_per_sec=njev / tot_t,
              cpu_baseline=dict(value=value, unit=UNIT, cores=1, kind="port",
                                sample=f"{args.workload}: first {nf} of {scene['F']} frames ({n} corners), scipy TRF+LSMR with 2-point FD Jacobian, "
                                       f"host has {os.cpu_count()} cores, numpy/scipy path is single threaded"),
              e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  emit(line)


def run_ours(args):
  import torch
  import torch.distributed as dist
  from multical_b200 import synthetic
  from multical_b200.calibration import from_scene, get_engine
  from multical_b200 import distributed as mdist

  rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
  assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
  torch.cuda.set_device(local)
  os.environ["MCBA_DEVICE"] = str(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

  # weak scaling: every rank holds the workload's frame count; cameras/boards (shared parameters) are common
  base = dict(synthetic.WORKLOADS[args.workload])
  scene = synthetic.make_scene(seed=args.seed, **{**base, "F": base["F"] * world})
  F_total = scene["F"]
  my_frames = mdist.frame_range(F_total, rank, world)
  local_scene = subsample_frames(scene, np.arange(*my_frames)) if world > 1 else scene
  # the step's inputs live in pinned host memory (the e2e timed region copies them to the device every step)
  for key in ("points", "valid"):
    pinned = torch.from_numpy(np.ascontiguousarray(local_scene[key])).pin_memory()
    local_scene[key] = pinned.numpy()
  calib = from_scene(local_scene).enable(cameras=True)
  eng = get_engine(local)
  stream = torch.cuda.current_stream()
  eng.lib.mcba_set_stream(eng.h, stream.cuda_stream)
  if world > 1:
    mdist.init_comm(eng, rank, world)
  n_local = int(calib.inliers.sum())
  n_total = n_local
  if world > 1:
    t = torch.tensor([n_local], dtype=torch.int64, device="cuda"); dist.all_reduce(t); n_total = int(t.item())

  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
  state0 = calib._state_arrays()

  def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident solves ------------------------------------------------------------------
  calib._upload(calib.inliers)
  def solve_resident():
    eng.set_params(*state0)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    res = eng.solve(ftol=BA_KW["tolerance"], max_nfev=BA_KW["max_iterations"], loss=BA_KW["loss"], f_scale=BA_KW["f_scale"])
    e1.record(stream); e1.synchronize()
    return res, e0.elapsed_time(e1)
  for _ in range(args.warmup): solve_resident()
  sampler = ClockSampler(local); sampler.start()
  barrier()
  t_dev, evals, njev, launches = 0.0, 0, 0, 0
  for _ in range(args.steps):
    res, ms = solve_resident()
    t_dev += ms; evals += res.nfev + res.njev; njev += res.njev; launches += res.kernel_launches
  barrier()
  tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
  if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  t_dev = float(tt.item())
  value = n_total * evals / (t_dev * 1e-3)

  # ---- end to end through the public API ---------------------------------------------------------
  def solve_e2e():
    c = from_scene(local_scene).enable(cameras=True)      # fresh object: nothing cached on host or device
    t0 = time.perf_counter()
    out = c.bundle_adjust(**BA_KW)          # returns after the device->host read of the solved parameter vector
    _ = out.last_solve.cost
    torch.cuda.synchronize()
    return out.last_solve, time.perf_counter() - t0
  for _ in range(args.warmup): solve_e2e()
  barrier()
  t_e2e, evals_e = 0.0, 0
  for _ in range(args.steps):
    flush.zero_(); torch.cuda.synchronize()
    res, dt = solve_e2e(); t_e2e += dt; evals_e += res.nfev + res.njev
  barrier()
  tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
  if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  t_e2e = float(tt.item())
  e2e_value = n_total * evals_e / t_e2e
  sampler.stop_flag = True; sampler.join(timeout=2)
  n_params = eng.num_params
  # dense upload: mask (1 B/entry) + observations (16 B/entry) of the [C,F,B,P] table, board points, parameter state
  h2d = int(calib.inliers.size) * (1 + 16) + int(np.prod(calib.board_points.points.shape)) * 8 + sum(a.size for a in state0) * 8
  d2h = n_params * 8 + 64

  if rank != 0:
    if world > 1: dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel (per-view moment accumulation, k_views_mma) ----------------
  NO_PREPARE = 256

  def time_moments_kernel(iters=20):
    info = eng.bench_info(0)
    eng.bench_launch(0, 3)                      # builds the pose tables, warms up
    times = []
    for _ in range(iters):
      flush.zero_()                             # inputs are smaller than L2 at cfg2: evict them between timed launches
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream); eng.bench_launch(0 | NO_PREPARE, 1); e1.record(stream); e1.synchronize()
      times.append(e0.elapsed_time(e1) * 1e-3 / info["launches_per_call"])
    return info, float(np.mean(times))

  calib._upload(calib.inliers)
  info, dur = time_moments_kernel()
  peak, which = measured_peak()
  achieved = info["bytes_per_launch"] / dur / 1e9
  roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=ncu_traffic(args.workload),
                  kernel="k_views_mma (per-view moment SYRK on the fp64 tensor path, %d launch per linearisation)" % info["launches_per_call"],
                  bytes_per_launch=info["bytes_per_launch"], launch_ms=dur * 1e3, peak_source=which,
                  note="fp64-pipe bound, not HBM bound: ~160 DFMA + 12 DMMA(m8n8k4) per corner against 18 B; ncu at 5.5M corners: "
                       "fp64+DMMA shared pipe 70% active, DRAM 5% (profiles/); below ~1M corners launch latency dominates")
  if world == 1 and args.at_scale:
    # the same kernel where it is not launch-latency bound: BASELINE configs[3] (16 cam x 1000 frames x 5 boards, ~5.5M corners)
    big = from_scene(synthetic.make_workload("cfg4", seed=args.seed)).enable(cameras=True)
    big._upload(big.inliers)
    binfo, bdur = time_moments_kernel(10)
    bach = binfo["bytes_per_launch"] / bdur / 1e9
    roofline["at_scale"] = dict(workload="cfg4", corners=binfo["corners"], bytes_per_launch=binfo["bytes_per_launch"], launch_ms=bdur * 1e3,
                                achieved=bach, frac=bach / peak)
    calib._upload(calib.inliers)

  # ---- CPU baseline: the oracle port on a bounded sample of the same workload --------------------
  nf = min(local_scene["F"], args.ref_frames)
  cpu_v, cpu_dt, cpu_res, cpu_n = cpu_reference_step(subsample_frames(local_scene, np.arange(nf)))
  cpu_baseline = dict(value=cpu_v, unit=UNIT, cores=1, kind="port",
                      sample=f"{args.workload}: first {nf} of {local_scene['F']} frames ({cpu_n} corners), one full bundle_adjust "
                             f"({cpu_res.nfev} nfev, {cpu_dt:.1f} s); host has {os.cpu_count()} cores, scipy/numpy path single threaded")

  line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
              ms_per_step=t_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
              data="synthetic",
              config=dict(workload=args.workload, cameras=scene["C"], frames=F_total, boards=scene["B"], corners=n_total,
                          params=n_params if world == 1 else None, frames_per_gpu=base["F"], camera_model=scene["model"],
                          solver="TRF semantics (ftol=1e-4, x_scale=jac, max_nfev=100), exact Schur inner solve",
                          l2="flushed between timed iterations (256 MiB write)", seed=args.seed,
                          kernel_variant=(" ".join(f"{k}={os.environ[k]}" for k in ("MCBA_MOMENTS", "MCBA_CHOL", "MCBA_FUSE", "MCBA_EXPAND", "MCBA_PEER") if k in os.environ) or "default")),
              lm_iters_per_sec=njev / (t_dev * 1e-3), nfev_plus_njev_per_step=evals / args.steps,
              e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=1e3 * t_e2e / args.steps,
                       lm_iters_per_sec=njev / t_e2e),
              gpu_launches=launches, clocks=sampler.summary(), roofline=roofline, cpu_baseline=cpu_baseline)
  emit(line)
  if world > 1: dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
  """The one JSON line goes to the real stdout; everything else this process (or NCCL) prints was sent to stderr."""
  out = os.fdopen(os.dup(_REAL_STDOUT), "w") if _REAL_STDOUT is not None else sys.stdout
  out.write(json.dumps(line) + "\n"); out.flush()


def main():
  global _REAL_STDOUT
  # libraries (NCCL's version banner, torchrun warnings) may write to fd 1: keep stdout clean for the single JSON line
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="cfg2")
  ap.add_argument("--seed", type=int, default=0)
  ap.add_argument("--ref-frames", type=int, default=20, help="frames in the CPU-baseline sample")
  ap.add_argument("--no-at-scale", dest="at_scale", action="store_false", help="skip the extra roofline point at cfg4 size")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    import __graft_entry__ as g
    g.build()
    run_ours(args)


if __name__ == "__main__":
  main()
