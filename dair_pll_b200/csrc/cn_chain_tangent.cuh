// Forward-mode derivative of the learnable time step of the generic serial chain (cn_chain.cuh instantiated on dual numbers,
// as cn_elbow_tangent.cuh does for the two-body tree): the backward of a chain rollout for the prediction loss
// (experiment.py:230-248; multibody_learnable_system.py:293-304).  Directions for N links, NX = 13 + 2 (N - 1):
//   0 .. 10N-1 inertia | 10N .. 11N-1 mu_pair | 11N .. 14N-1 half | 14N .. 14N+NX-1 the coordinates of x0.
#pragma once
#include "cn_chain.cuh"
#include "cn_dual.cuh"

namespace cn {

// xbar: upstream gradient w.r.t. traj[1..steps] (steps x NX).  Returns sum_s xbar_s . d x_s / d (direction).
template <typename B, int N>
CN_HD B chain_rollout_tangent(const B* inertia, const B* mu, const B* half, const B* kin, B dt, B eps, const B* x0, int steps,
                              const B* xbar, int dir) {
  typedef DualN<B, 1> D;
  constexpr int NX = 13 + 2 * (N - 1);
  D din[10 * N], dmu[N], dh[3 * N], dkin[CH_NKIN * N];
  for (int i = 0; i < 10 * N; ++i) { din[i] = D(inertia[i]); if (dir == i) din[i].d[0] = B(1); }
  for (int i = 0; i < N; ++i) { dmu[i] = D(mu[i]); if (dir == 10 * N + i) dmu[i].d[0] = B(1); }
  for (int i = 0; i < 3 * N; ++i) { dh[i] = D(half[i]); if (dir == 11 * N + i) dh[i].d[0] = B(1); }
  for (int i = 0; i < CH_NKIN * N; ++i) dkin[i] = D(kin[i]);
  ChainParams<D, N> P;
  chain_params_init<D, N>(P, din, dmu, dh, dkin, D(dt), D(eps));
  const SolverCfg<B> c0 = default_cfg<B>();
  SolverCfg<D> cfg;
  cfg.tol_rel = D(c0.tol_rel); cfg.tol_stall = D(c0.tol_stall); cfg.ls_c = D(c0.ls_c); cfg.max_iter = c0.max_iter;
  cfg.tol_final = D(0);      // every solve ends with the polishing step at the converged point, whose tangent is the
  cfg.polish = true;         // implicit-function derivative of the QP solution
  D x[NX], xn[NX];
  for (int i = 0; i < NX; ++i) { x[i] = D(x0[i]); if (dir == 14 * N + i) x[i].d[0] = B(1); }
  B g = B(0);
  for (int s = 0; s < steps; ++s) {
    chain_step_sample<D, N>(P, cfg, x, xn);
    for (int i = 0; i < NX; ++i) { g += xbar[s * NX + i] * xn[i].d[0]; x[i] = xn[i]; }
  }
  return g;
}

}  // namespace cn
