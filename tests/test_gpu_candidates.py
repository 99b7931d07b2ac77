"""GPU (-m gpu): opt-in kernel variants that wait for an A/B measurement must at least be exact replacements.

MCBA_CHOL=blocked -- k_chol_blocked, the single-CTA blocked reduced solve with warp-level column steps, instead of k_chol_small: the
iteration table of a solve must agree with the default kernel's to round-off."""
import numpy as np
import pytest

import test_gpu_parity as gp
from multical_b200 import calibration

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["standard_2x6", "cube3_3x6", "fisheye_3x5"])
def test_blocked_reduced_solve_reproduces_the_default_iterations(name, monkeypatch):
  def solve():
    scene, z, calib, prob = gp.make(name)
    return calib.bundle_adjust(tolerance=1e-12, max_iterations=40).last_solve
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
