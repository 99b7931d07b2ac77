"""GPU (-m gpu): opt-in kernel variants that wait for an A/B measurement must at least be exact replacements.

MCBA_CHOL=blocked -- k_chol_blocked, the single-CTA blocked reduced solve with warp-level column steps, instead of k_chol_small: the
iteration table of a solve must agree with the default kernel's to round-off.
MCBA_FUSE=1 -- four launches fewer per LM iteration (k_dots folded into the second k_quad, k_step + k_make_trial as one single-CTA
launch, k_accept / k_scale as tails of the moment kernel / k_expand_shared): same iterations up to the summation order of a few sums.
MCBA_MOMENTS=f32 -- k_views_f32: Hessian moments in FP32, residual / cost / gradient in FP64.  Not an exact replacement by design:
the iteration path may differ, the minimiser may not -- converged cost within 1e-10 relative, evaluations within +-2, and the
parity hook mcba_linearize keeps returning the fp64 normal equations.
MCBA_EXPAND=parallel -- the per-view twist maps of the expand kernels computed by the whole warp in three short phases instead of by
3 (5) lanes with ~300 dependent flops each: the same products in the same order, so the normal equations must be IDENTICAL."""
import numpy as np
import pytest

import test_gpu_parity as gp
from multical_b200 import calibration

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6", "poses_only_2x6"])
def test_fused_launches_reproduce_the_default_iterations(name, monkeypatch):
  def solve():
    scene, z, calib, prob = gp.make(name)
    return calib.bundle_adjust(tolerance=1e-9, max_iterations=40).last_solve      # away from ties between the ftol and xtol tests
  ref = solve()
  monkeypatch.setenv("MCBA_FUSE", "1")                  # read by mcba_create: a fresh context is needed
  for eng in calibration._engines.values(): eng.close()
  monkeypatch.setattr(calibration, "_engines", {})
  got = solve()
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  assert got.nfev == ref.nfev and got.status == ref.status
  assert got.kernel_launches < ref.kernel_launches
  assert np.allclose(np.array(ref.log, float)[:, 2], np.array(got.log, float)[:, 2], rtol=1e-9, atol=0)


@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6", "fisheye_3x5"])
def test_blocked_reduced_solve_reproduces_the_default_iterations(name, monkeypatch):
  def solve():
    scene, z, calib, prob = gp.make(name)
    return calib.bundle_adjust(tolerance=1e-9, max_iterations=40).last_solve      # away from ties between the ftol and xtol tests
  ref = solve()
  monkeypatch.setenv("MCBA_CHOL", "blocked")            # read by mcba_create: a fresh context is needed
  for eng in calibration._engines.values(): eng.close()
  monkeypatch.setattr(calibration, "_engines", {})
  got = solve()
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  assert got.nfev == ref.nfev and got.status == ref.status and got.chol_retries == 0
  assert abs(got.cost - ref.cost) <= 1e-12 * ref.cost
  a, b = np.array(ref.log, float), np.array(got.log, float)
  assert a.shape == b.shape
  assert np.allclose(a[:, 2], b[:, 2], rtol=1e-10, atol=0)          # cost column of every iteration


@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6", "fisheye_3x5", "invalid_poses_3x6"])
def test_fp32_hessian_moments_reach_the_same_minimum(name, monkeypatch):
  def solve(tol):
    scene, z, calib, prob = gp.make(name)
    eng = calib._upload(calib.inliers)
    JtJ, Jtr, cost = eng.linearize(z["x1"])
    return calib.bundle_adjust(tolerance=tol, max_iterations=60).last_solve, JtJ
  ref, H = solve(1e-12)
  ref4, _ = solve(1e-4)
  monkeypatch.setenv("MCBA_MOMENTS", "f32")             # read by mcba_create: a fresh context is needed
  for eng in calibration._engines.values(): eng.close()
  monkeypatch.setattr(calibration, "_engines", {})
  got, H32 = solve(1e-12)
  got4, _ = solve(1e-4)
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  assert np.array_equal(H, H32)                                               # the hook is not affected
  assert abs(got.cost - ref.cost) <= 1e-10 * ref.cost and abs(got.nfev - ref.nfev) <= 2
  assert abs(got4.cost - ref4.cost) <= 1e-6 * ref4.cost and abs(got4.nfev - ref4.nfev) <= 2      # the reference's default tolerance


@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6", "invalid_poses_3x6", "tilted_2x5"])
def test_warp_parallel_twist_maps_give_identical_normal_equations(name, monkeypatch):
  def linearise():
    scene, z, calib, prob = gp.make(name)
    return calib._upload(calib.inliers).linearize(z["x1"])
  H, g, c = linearise()
  monkeypatch.setenv("MCBA_EXPAND", "parallel")         # read by mcba_create: a fresh context is needed
  for eng in calibration._engines.values(): eng.close()
  monkeypatch.setattr(calibration, "_engines", {})
  H2, g2, c2 = linearise()
  for eng in calibration._engines.values(): eng.close()
  calibration._engines.clear()
  assert np.abs(H - H2).max() <= 1e-13 * np.abs(H).max() and np.abs(g - g2).max() <= 1e-13 * np.abs(g).max() and c == c2
