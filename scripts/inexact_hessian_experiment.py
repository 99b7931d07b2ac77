"""CPU experiment behind the MCBA_MOMENTS=f32 candidate: scipy's trust-region iteration (oracle/trf_exact_model.py logic)
with the EXACT gradient J^T f but the Hessian model J^T J formed from a Jacobian rounded to a given number of mantissa bits
(24 = fp32, 21 ~ 3xTF32, 11 = TF32, 8 = bf16).  Prints final cost / number of evaluations / status per precision and tolerance.
Uses oracle/ (test infrastructure): run it from the repo root; not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from numpy.linalg import norm
from scipy.optimize._numdiff import approx_derivative, group_columns
from scipy.optimize._lsq.common import check_termination, minimize_quadratic_1d, solve_trust_region_2d, update_tr_radius
from multical_b200 import synthetic
from oracle.ba_oracle import Problem

def rnd(J, bits):
  if bits is None: return J
  m, e = np.frexp(J)
  return np.ldexp(np.round(m * 2.0 ** bits) / 2.0 ** bits, e)

def trf(fun, jac, x0, bits, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=60):
  x = np.array(x0, float); f = fun(x); nfev = 1; J = jac(x); cost = 0.5 * f @ f; g = J.T @ f
  Jt = rnd(J, bits)
  scale_inv = np.sqrt((Jt ** 2).sum(0)); scale_inv[scale_inv == 0] = 1
  Delta = norm(x * scale_inv) or 1.0
  status = None; costs = [cost]
  while True:
    if norm(g, np.inf) < gtol: status = 1
    if status is not None or nfev >= max_nfev: break
    d = 1.0 / scale_inv; g_h = d * g; J_h = Jt * d
    A = J_h.T @ J_h
    a, b = g_h @ A @ g_h, -(g_h @ g_h)
    reg = max(-minimize_quadratic_1d(a, b, 0, Delta / norm(g_h))[1] / Delta ** 2, 1e-12)
    gn_h = np.linalg.solve(A + reg * np.eye(A.shape[0]), g_h)
    S, _ = np.linalg.qr(np.vstack((g_h, gn_h)).T)
    JS = J_h @ S; B_S, g_S = JS.T @ JS, S.T @ g_h
    reduction = -1
    while reduction <= 0 and nfev < max_nfev:
      p_S, _ = solve_trust_region_2d(B_S, g_S, Delta)
      step_h = S @ p_S; Js = J_h @ step_h
      predicted = -(0.5 * Js @ Js + g_h @ step_h)
      step = d * step_h
      f_new = fun(x + step); nfev += 1
      cost_new = 0.5 * f_new @ f_new; reduction = cost - cost_new
      shn = norm(step_h)
      Delta_new, ratio = update_tr_radius(Delta, reduction, predicted, shn, shn > 0.95 * Delta)
      status = check_termination(reduction, cost, norm(step), norm(x), ratio, ftol, xtol)
      if status is not None: break
      Delta = Delta_new
    if reduction > 0:
      x = x + step; f = f_new; cost = cost_new; costs.append(cost)
      J = jac(x); g = J.T @ f; Jt = rnd(J, bits)
      scale_inv = np.maximum(scale_inv, np.sqrt((Jt ** 2).sum(0)))
  return cost, nfev, status, costs

for kw in [dict(C=2, F=6, vis=0.5, seed=11), dict(C=3, F=6, vis=0.6, seed=14, boards=("cube", 10, 10, 0.04, 3), rig="dome"), dict(C=4, F=20, vis=0.3, seed=5),
           dict(C=3, F=10, vis=0.4, seed=7, model="fisheye")]:
  scene = synthetic.make_scene(**kw)
  prob = Problem.from_scene(scene, optimize=dict(cameras=True))
  S = prob.sparsity_matrix(); groups = group_columns(S)
  jac = lambda x: approx_derivative(prob.residuals, x, method="3-point", sparsity=(S, groups)).toarray()
  print(kw)
  for name, bits in [("fp64", None), ("fp32", 24), ("tf32x2", 21), ("tf32", 11), ("bf16", 8)]:
    for tol in (1e-4, 1e-10):
      cost, nfev, status, costs = trf(prob.residuals, jac, prob.param_vec, bits, ftol=tol)
      print(f"  {name:7s} ftol {tol:g}: cost {cost:.9f} nfev {nfev} status {status}")
