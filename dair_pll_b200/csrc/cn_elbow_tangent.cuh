// Forward-mode derivative of the learnable time step of the two-body (elbow) system: the step code of
// cn_elbow.cuh instantiated on dual numbers carrying one tangent, as cn_cube_tangent.cuh does for the
// cube.  43 directions:
//   0..27  callable-level parameters [inertia 2x10 | mu_pair 2 | half 2x3],   28..42  the 15 coordinates of x0.
// Box geometries only (witness points of a learned geometry depend on the state through the caller's
// networks, so that rollout is stepped on the host).
#pragma once
#include "cn_elbow.cuh"
#include "cn_dual.cuh"
#include "cn_elbow_wf.cuh"

namespace cn {

constexpr int ELBOW_NTAN = 43;
constexpr int ELBOW_NPARAM_TAN = 28;

// xbar: upstream gradient w.r.t. traj[1..steps] (steps x 15).  Returns sum_s xbar_s . d x_s / d (direction).
template <typename B>
CN_HD B elbow_rollout_tangent(const B* inertia, const B* mu, const B* half, const B* kin, B dt, B eps, const B* x0,
                              int steps, const B* xbar, int dir, const B* usol = nullptr) {
  typedef DualN<B, 1> D;
  D din[20], dmu[2], dh[6], dkin[12];
  for (int i = 0; i < 20; ++i) { din[i] = D(inertia[i]); if (dir == i) din[i].d[0] = B(1); }
  for (int i = 0; i < 2; ++i) { dmu[i] = D(mu[i]); if (dir == 20 + i) dmu[i].d[0] = B(1); }
  for (int i = 0; i < 6; ++i) { dh[i] = D(half[i]); if (dir == 22 + i) dh[i].d[0] = B(1); }
  for (int i = 0; i < 12; ++i) dkin[i] = D(kin[i]);
  ElbowParams<D> P;
  elbow_params_init<D>(P, din, dmu, dh, dkin, D(dt), D(eps));
  const SolverCfg<B> c0 = default_cfg<B>();
  SolverCfg<D> cfg;
  cfg.tol_rel = D(c0.tol_rel); cfg.tol_stall = D(c0.tol_stall); cfg.ls_c = D(c0.ls_c); cfg.max_iter = c0.max_iter;
  cfg.tol_final = D(0);      // no early finish: every solve ends with the polishing step at the converged point,
  cfg.polish = true;         // whose tangent is the exact implicit derivative of the QP solution
  D x[15], xn[15];
  for (int i = 0; i < 15; ++i) { x[i] = D(x0[i]); if (dir == ELBOW_NPARAM_TAN + i) x[i].d[0] = B(1); }
  B g = B(0);
  for (int s = 0; s < steps; ++s) {
    if (usol) {      // the forward rollout kept every step's optimum: one Newton step at it instead of a dual-number solve
      D uf[7];
      for (int i = 0; i < 7; ++i) uf[i] = D(usol[s * 7 + i]);
      elbow_step_sample_wf<D>(P, cfg, x, (const D*)nullptr, xn, (D*)nullptr, (D*)nullptr, uf);
    } else {
      elbow_step_sample<D>(P, cfg, x, (const D*)nullptr, xn, (D*)nullptr);
    }
    for (int i = 0; i < 15; ++i) { g += xbar[s * 15 + i] * xn[i].d[0]; x[i] = xn[i]; }
  }
  return g;
}

// One learnable step with caller-supplied witness points (learned geometry): 61 directions
//   0..19 inertia | 20..21 mu_pair | 22..45 the 8 x 3 witness-point coordinates | 46..60 the 15 coordinates of x.
// Returns xbar . d x_next / d (direction).  The caller's networks make the points a function of the state and of their
// weights; a piecewise-linear support function has piecewise-CONSTANT support points, so the chain through the
// directions vanishes almost everywhere and the points' cotangent goes to the weights alone (ops.ElbowStepPts).
constexpr int ELBOW_PTS_NTAN = 61;
constexpr int ELBOW_PTS_NPARAM = 22;

template <typename B>
CN_HD B elbow_step_pts_tangent(const B* inertia, const B* mu, const B* kin, B dt, B eps, const B* x0, const B* pts,
                               const B* xbar, int dir, const B* usol = nullptr) {
  typedef DualN<B, 1> D;
  D din[20], dmu[2], dh[6], dkin[12], dp[24];
  for (int i = 0; i < 20; ++i) { din[i] = D(inertia[i]); if (dir == i) din[i].d[0] = B(1); }
  for (int i = 0; i < 2; ++i) { dmu[i] = D(mu[i]); if (dir == 20 + i) dmu[i].d[0] = B(1); }
  for (int i = 0; i < 6; ++i) dh[i] = D(B(0));
  for (int i = 0; i < 12; ++i) dkin[i] = D(kin[i]);
  for (int i = 0; i < 24; ++i) { dp[i] = D(pts[i]); if (dir == ELBOW_PTS_NPARAM + i) dp[i].d[0] = B(1); }
  ElbowParams<D> P;
  elbow_params_init<D>(P, din, dmu, dh, dkin, D(dt), D(eps));
  const SolverCfg<B> c0 = default_cfg<B>();
  SolverCfg<D> cfg;
  cfg.tol_rel = D(c0.tol_rel); cfg.tol_stall = D(c0.tol_stall); cfg.ls_c = D(c0.ls_c); cfg.max_iter = c0.max_iter;
  cfg.tol_final = D(0);
  cfg.polish = true;
  D x[15], xn[15];
  for (int i = 0; i < 15; ++i) { x[i] = D(x0[i]); if (dir == ELBOW_PTS_NPARAM + 24 + i) x[i].d[0] = B(1); }
  if (usol) {
    D uf[7];
    for (int i = 0; i < 7; ++i) uf[i] = D(usol[i]);
    elbow_step_sample_wf<D>(P, cfg, x, dp, xn, (D*)nullptr, (D*)nullptr, uf);
  } else {
    elbow_step_sample<D>(P, cfg, x, dp, xn, (D*)nullptr);
  }
  B g = B(0);
  for (int i = 0; i < 15; ++i) g += xbar[i] * xn[i].d[0];
  return g;
}

}  // namespace cn
