// Forward-mode derivative of the learnable time step (the backward of forward_dynamics / sim_step /
// VelocityIntegrator.step that the reference gets from autograd through sappy,
// multibody_learnable_system.py:293-304; integrator.py:153-162): the step code of cn_cube.cuh is
// instantiated on dual numbers (cn_dual.cuh) carrying ONE tangent, and the rollout is run once per
// direction -- 27 directions:
//   0..13  callable-level parameters [inertia 10 | mu_pair 1 | half 3],   14..26  the 13 coordinates of x0 --
// contracting the tangent of every x_s with the upstream gradient on the fly.  One (sample,
// direction) pair per thread: 27x the forward work, but no tape, constant memory in the number of
// steps, and parallelism even for the small batches of the prediction loss.
#pragma once
#include "cn_cube.cuh"
#include "cn_dual.cuh"

namespace cn {

constexpr int CUBE_NTAN = 27;

// xbar: upstream gradient w.r.t. traj[1..steps] (steps x 13, row s = step s+1).
// Returns sum_s xbar_s . d x_s / d (direction `dir`).
template <typename B>
CN_HD B cube_rollout_tangent(const B* inertia, const B* mu, const B* half, B dt, B eps, const B* x0, int steps,
                             const B* xbar, int dir) {
  typedef DualN<B, 1> D;
  D din[10], dmu[1], dh[3];
  for (int i = 0; i < 10; ++i) { din[i] = D(inertia[i]); if (dir == i) din[i].d[0] = B(1); }
  dmu[0] = D(mu[0]); if (dir == 10) dmu[0].d[0] = B(1);
  for (int i = 0; i < 3; ++i) { dh[i] = D(half[i]); if (dir == 11 + i) dh[i].d[0] = B(1); }
  CubeParams<D> P;
  cube_params_init<D>(P, din, dmu, dh, D(dt), D(eps));
  const SolverCfg<B> c0 = default_cfg<B>();
  SolverCfg<D> cfg;
  cfg.tol_rel = D(c0.tol_rel); cfg.tol_stall = D(c0.tol_stall); cfg.ls_c = D(c0.ls_c); cfg.max_iter = c0.max_iter;
  cfg.tol_final = D(0);      // no early finish: every solve ends with the polishing step at the converged point,
  cfg.polish = true;         // whose tangent is the exact implicit derivative of the QP solution
  D x[13], xn[13];
  for (int i = 0; i < 13; ++i) { x[i] = D(x0[i]); if (dir == 14 + i) x[i].d[0] = B(1); }
  B g = B(0);
  for (int s = 0; s < steps; ++s) {
    cube_step_sample<D>(P, cfg, x, xn, (D*)nullptr);
    for (int i = 0; i < 13; ++i) { g += xbar[s * 13 + i] * xn[i].d[0]; x[i] = xn[i]; }
  }
  return g;
}

// One learnable step of a single floating body with caller-supplied witness points (Sphere, Polygon: geometry.py:415-456,
// 220-252 through collide_plane_convex :553-582): 36 directions
//   0..9 inertia | 10 mu_pair | 11..22 the 4 x 3 witness-point coordinates | 23..35 the 13 coordinates of x.
// Returns xbar . d x_next / d (direction).  The points' cotangent is chained by the caller's autograd into the shape
// parameters AND into the state (a sphere's support point d r moves with the orientation).
constexpr int BODY_PTS_NTAN = 36;

template <typename B>
CN_HD B body_step_pts_tangent(const B* inertia, const B* mu, B dt, B eps, const B* x0, const B* pts, int n_c,
                              const B* xbar, int dir) {
  typedef DualN<B, 1> D;
  D din[10], dmu[1], dh[3], dp[12];
  for (int i = 0; i < 10; ++i) { din[i] = D(inertia[i]); if (dir == i) din[i].d[0] = B(1); }
  dmu[0] = D(mu[0]); if (dir == 10) dmu[0].d[0] = B(1);
  for (int i = 0; i < 3; ++i) dh[i] = D(B(0));
  for (int i = 0; i < 12; ++i) { dp[i] = D(pts[i]); if (dir == 11 + i) dp[i].d[0] = B(1); }
  CubeParams<D> P;
  cube_params_init<D>(P, din, dmu, dh, D(dt), D(eps));
  const SolverCfg<B> c0 = default_cfg<B>();
  SolverCfg<D> cfg;
  cfg.tol_rel = D(c0.tol_rel); cfg.tol_stall = D(c0.tol_stall); cfg.ls_c = D(c0.ls_c); cfg.max_iter = c0.max_iter;
  cfg.tol_final = D(0);
  cfg.polish = true;
  D x[13], xn[13];
  for (int i = 0; i < 13; ++i) { x[i] = D(x0[i]); if (dir == 23 + i) x[i].d[0] = B(1); }
  body_step_sample_pts<D>(P, cfg, x, dp, n_c, xn, (D*)nullptr);
  B g = B(0);
  for (int i = 0; i < 13; ++i) g += xbar[i] * xn[i].d[0];
  return g;
}

}  // namespace cn
