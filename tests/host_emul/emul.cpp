// Host build of the per-sample device math (dair_pll_b200/csrc/*.cuh) for debugging in a
// container without a GPU.  TEST INFRASTRUCTURE ONLY: nothing in the package loads this.
#include "../../dair_pll_b200/csrc/cn_cube.cuh"
#include "../../dair_pll_b200/csrc/cn_params.cuh"
#include "../../dair_pll_b200/csrc/cn_elbow.cuh"
#include "../../dair_pll_b200/csrc/cn_cube_tangent.cuh"
#include "../../dair_pll_b200/csrc/cn_elbow_tangent.cuh"
#include "../../dair_pll_b200/csrc/cn_cube_adjoint.cuh"
#include "../../dair_pll_b200/csrc/cn_elbow_wf.cuh"
#include "../../dair_pll_b200/csrc/cn_chain.cuh"
#include "../../dair_pll_b200/csrc/cn_icnn_tc.cuh"
#include <vector>
#include <cstdint>
using namespace cn;
// generic kinematic tree (cn_chain.cuh), N = 2 .. 6 links
template <int N>
static int chain_loss_emul(const double* x, const double* xp, const double* inertia, const double* mu, const double* half,
                           const double* kin, double dt, double eps, int64_t B, double* loss, double* force, int32_t* iters,
                           double* grad) {
  ChainParams<double, N> P;
  chain_params_init<double, N>(P, inertia, mu, half, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  constexpr int NX = 13 + 2 * (N - 1);
  for (int i = 0; i < 14 * N; ++i) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = chain_loss_sample<double, N>(P, cfg, x + NX * b, xp + NX * b, grad, force ? force + 12 * N * b : nullptr, &it);
    if (iters) iters[b] = it;
  }
  return 0;
}
template <int N>
static int chain_step_emul(const double* x, const double* inertia, const double* mu, const double* half, const double* kin,
                           double dt, double eps, int64_t B, double* xn) {
  ChainParams<double, N> P;
  chain_params_init<double, N>(P, inertia, mu, half, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  constexpr int NX = 13 + 2 * (N - 1);
  for (int64_t b = 0; b < B; ++b) chain_step_sample<double, N>(P, cfg, x + NX * b, xn + NX * b);
  return 0;
}
template <int N>
static int chain_terms_emul(int g, const double* q, const double* v, const double* inertia, const double* mu, const double* half,
                            const double* kin, int64_t B, double* M, double* J, double* phi, double* acc, double* D) {
  constexpr int NV = 6 + N - 1, NQ = 7 + N - 1;
  ChainParams<double, N> P;
  chain_params_init<double, N>(P, inertia, mu, half, kin, 1.0, 1.0);
  const int nc = 4 * g, kk = 3 * nc;
  for (int64_t b = 0; b < B; ++b)
    chain_terms_sample<double, N>(P, q + NQ * b, v + NV * b, g, M + NV * NV * b, J + kk * NV * b, phi + nc * b, acc + NV * b,
                                  D + kk * kk * b);
  return 0;
}

template <int N>
static int chain_loss_pts_emul(const double* x, const double* xp, const double* inertia, const double* mu, const double* kin,
                               const double* pts, unsigned npts, double dt, double eps, int64_t B, double* loss, double* grad,
                               double* grad_pts) {
  constexpr int NX = 13 + 2 * (N - 1);
  double half[3 * N] = {0};
  ChainParams<double, N> P;
  chain_params_init<double, N>(P, inertia, mu, half, kin, dt, eps);
  const SolverCfg<double> cfg = default_cfg<double>();
  for (int i = 0; i < 14 * N; ++i) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = chain_loss_sample<double, N>(P, cfg, x + NX * b, xp + NX * b, grad, nullptr, &it, pts + 12 * N * b, npts,
                                           grad_pts + 12 * N * b);
  }
  return 0;
}
template <int N>
static int chain_step_pts_emul(const double* x, const double* inertia, const double* mu, const double* kin, const double* pts,
                               unsigned npts, double dt, double eps, int64_t B, double* xn) {
  constexpr int NX = 13 + 2 * (N - 1);
  double half[3 * N] = {0};
  ChainParams<double, N> P;
  chain_params_init<double, N>(P, inertia, mu, half, kin, dt, eps);
  const SolverCfg<double> cfg = default_cfg<double>();
  for (int64_t b = 0; b < B; ++b) chain_step_sample<double, N>(P, cfg, x + NX * b, xn + NX * b, pts + 12 * N * b, npts);
  return 0;
}

extern "C" {
// digit planes of the tensor-core support-point kernel (cn_icnn_tc.cuh): column of n weights -> digits, value the planes
// stand for, image offsets
int emul_tc_digits(const double* q, int n, int8_t* dig, double* rec, double* sigma_out) {
  double cmax = 0;
  for (int i = 0; i < n; ++i) cmax = std::fmax(cmax, std::fabs(q[i]));
  int e;
  *sigma_out = cn::tc_column_scale(cmax, &e);
  for (int i = 0; i < n; ++i) {
    cn::tc_digits(q[i], e, dig + cn::TC_NS * i);
    rec[i] = cn::tc_reconstruct(dig + cn::TC_NS * i, e);
  }
  return cn::TC_NS;
}
int emul_tc_image_offset(int k, int s, int j, int i) { return cn::tc_image_offset(k, s, j, i); }
int emul_tc_image_bytes() { return cn::TC_IMG_BYTES; }

int emul_cube_loss_f64(const double* x, const double* xp, const double* inertia, const double* mu,
                       const double* half, double dt, double eps, int64_t B, double* loss, double* force,
                       int32_t* iters, double* grad) {
  CubeParams<double> P;
  cube_params_init(P, inertia, mu, half, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int i = 0; i < CUBE_NPARAM; ++i) if (grad) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = cube_loss_sample(P, cfg, x + 13 * b, xp + 13 * b, grad, force ? force + 12 * b : nullptr, &it);
    if (iters) iters[b] = it;
  }
  return 0;
}
// free-flight fast path (triage phase of the wavefront kernel): flags[b] = 1 where it applies
int emul_cube_loss_free_f64(const double* x, const double* xp, const double* inertia, const double* mu,
                            const double* half, double dt, double eps, int64_t B, double* loss, int32_t* flags,
                            double* grad) {
  CubeParams<double> P;
  cube_params_init(P, inertia, mu, half, dt, eps);
  for (int i = 0; i < CUBE_NPARAM; ++i) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    double l = 0;
    flags[b] = cube_loss_free_flight<double>(P, x + 13 * b, xp + 13 * b, grad, &l) ? 1 : 0;
    loss[b] = l;
  }
  return 0;
}
// fp32 variant = fp32 storage, fp64 arithmetic (as the kernels: T = double, IO = float)
int emul_cube_loss_f32(const float* x, const float* xp, const float* inertia, const float* mu,
                       const float* half, float dt, float eps, int64_t B, float* loss, float* force,
                       int32_t* iters, float* grad) {
  double in[10], m[1], h[3];
  for (int i = 0; i < 10; ++i) in[i] = inertia[i];
  m[0] = mu[0];
  for (int i = 0; i < 3; ++i) h[i] = half[i];
  CubeParams<double> P;
  cube_params_init(P, in, m, h, (double)dt, (double)eps);
  SolverCfg<double> cfg = default_cfg<double>();
  double g[CUBE_NPARAM];
  for (int i = 0; i < CUBE_NPARAM; ++i) g[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    double xs[13], xps[13], fo[12];
    for (int i = 0; i < 13; ++i) { xs[i] = x[13 * b + i]; xps[i] = xp[13 * b + i]; }
    int it;
    loss[b] = (float)cube_loss_sample(P, cfg, xs, xps, grad ? g : nullptr, force ? fo : nullptr, &it);
    if (force) for (int i = 0; i < 12; ++i) force[12 * b + i] = (float)fo[i];
    if (iters) iters[b] = it;
  }
  if (grad) for (int i = 0; i < CUBE_NPARAM; ++i) grad[i] = (float)g[i];
  return 0;
}
int emul_cube_step_f64(const double* x, const double* inertia, const double* mu, const double* half,
                       double dt, double eps, int64_t B, double* xn, double* force, int32_t* iters) {
  CubeParams<double> P;
  cube_params_init(P, inertia, mu, half, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int64_t b = 0; b < B; ++b) {
    int it = cube_step_sample(P, cfg, x + 13 * b, xn + 13 * b, force ? force + 12 * b : nullptr);
    if (iters) iters[b] = it;
  }
  return 0;
}
int emul_cube_terms_f64(const double* q, const double* v, const double* inertia, const double* mu, const double* half,
                        int64_t B, double* M, double* J, double* phi, double* acc, double* D) {
  CubeParams<double> P;
  cube_params_init(P, inertia, mu, half, 1.0, 1.0);
  for (int64_t b = 0; b < B; ++b)
    cube_terms_sample<double>(P, q + 7 * b, v + 6 * b, M + 36 * b, J + 72 * b, phi + 4 * b, acc + 6 * b, D + 144 * b);
  return 0;
}
int emul_chain_loss_f64(int n, const double* x, const double* xp, const double* inertia, const double* mu,
                        const double* half, const double* kin, double dt, double eps, int64_t B, double* loss, double* force,
                        int32_t* iters, double* grad) {
  if (n == 2) return chain_loss_emul<2>(x, xp, inertia, mu, half, kin, dt, eps, B, loss, force, iters, grad);
  if (n == 3) return chain_loss_emul<3>(x, xp, inertia, mu, half, kin, dt, eps, B, loss, force, iters, grad);
  if (n == 4) return chain_loss_emul<4>(x, xp, inertia, mu, half, kin, dt, eps, B, loss, force, iters, grad);
  if (n == 5) return chain_loss_emul<5>(x, xp, inertia, mu, half, kin, dt, eps, B, loss, force, iters, grad);
  if (n == 6) return chain_loss_emul<6>(x, xp, inertia, mu, half, kin, dt, eps, B, loss, force, iters, grad);
  return -1;
}
int emul_chain_step_f64(int n, const double* x, const double* inertia, const double* mu, const double* half,
                        const double* kin, double dt, double eps, int64_t B, double* xn) {
  if (n == 2) return chain_step_emul<2>(x, inertia, mu, half, kin, dt, eps, B, xn);
  if (n == 3) return chain_step_emul<3>(x, inertia, mu, half, kin, dt, eps, B, xn);
  if (n == 4) return chain_step_emul<4>(x, inertia, mu, half, kin, dt, eps, B, xn);
  if (n == 5) return chain_step_emul<5>(x, inertia, mu, half, kin, dt, eps, B, xn);
  if (n == 6) return chain_step_emul<6>(x, inertia, mu, half, kin, dt, eps, B, xn);
  return -1;
}
// single floating body with witness points (Sphere / Polygon / any plane-convex pair)
int emul_body_loss_pts_f64(const double* x, const double* xp, const double* inertia, const double* mu, const double* pts,
                           int n_c, double dt, double eps, int64_t B, double* loss, double* grad11, double* grad_pts) {
  CubeParams<double> P;
  double h[3] = {0, 0, 0};
  cube_params_init(P, inertia, mu, h, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int i = 0; i < 11; ++i) grad11[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = body_loss_sample_pts<double>(P, cfg, x + 13 * b, xp + 13 * b, pts + 12 * b, n_c, grad11, grad_pts + 12 * b,
                                           (double*)nullptr, &it);
  }
  return 0;
}
int emul_body_step_pts_f64(const double* x, const double* inertia, const double* mu, const double* pts, int n_c, double dt,
                           double eps, int64_t B, double* xn) {
  CubeParams<double> P;
  double h[3] = {0, 0, 0};
  cube_params_init(P, inertia, mu, h, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int64_t b = 0; b < B; ++b)
    body_step_sample_pts<double>(P, cfg, x + 13 * b, pts + 12 * b, n_c, xn + 13 * b, (double*)nullptr);
  return 0;
}
int emul_elbow_terms_f64(const double* q, const double* v, const double* inertia, const double* mu, const double* half,
                         const double* kin, int64_t B, double* M, double* J, double* phi, double* acc, double* D) {
  ElbowParams<double> P;
  elbow_params_init(P, inertia, mu, half, kin, 1.0, 1.0);
  for (int64_t b = 0; b < B; ++b)
    elbow_terms_sample<double>(P, q + 8 * b, v + 7 * b, M + 49 * b, J + 168 * b, phi + 8 * b, acc + 7 * b, D + 576 * b);
  return 0;
}
int emul_chain_loss_pts_f64(int n, const double* x, const double* xp, const double* inertia, const double* mu,
                            const double* kin, const double* pts, unsigned npts, double dt, double eps, int64_t B,
                            double* loss, double* grad, double* grad_pts) {
  if (n == 3) return chain_loss_pts_emul<3>(x, xp, inertia, mu, kin, pts, npts, dt, eps, B, loss, grad, grad_pts);
  if (n == 4) return chain_loss_pts_emul<4>(x, xp, inertia, mu, kin, pts, npts, dt, eps, B, loss, grad, grad_pts);
  return 1;
}
int emul_chain_step_pts_f64(int n, const double* x, const double* inertia, const double* mu, const double* kin,
                            const double* pts, unsigned npts, double dt, double eps, int64_t B, double* xn) {
  if (n == 3) return chain_step_pts_emul<3>(x, inertia, mu, kin, pts, npts, dt, eps, B, xn);
  if (n == 4) return chain_step_pts_emul<4>(x, inertia, mu, kin, pts, npts, dt, eps, B, xn);
  return 1;
}
int emul_chain_terms_f64(int n, int g, const double* q, const double* v, const double* inertia, const double* mu,
                         const double* half, const double* kin, int64_t B, double* M, double* J, double* phi, double* acc,
                         double* D) {
  if (n == 3) return chain_terms_emul<3>(g, q, v, inertia, mu, half, kin, B, M, J, phi, acc, D);
  if (n == 4) return chain_terms_emul<4>(g, q, v, inertia, mu, half, kin, B, M, J, phi, acc, D);
  return 1;
}
// theta -> inertia vector and the reverse-direction product g^T J via dual numbers (as the reduce kernel does)
int emul_theta_chain_f64(const double* theta, const double* g_inertia, double* inertia, double* grad_theta) {
  double th[10], out[10];
  for (int i = 0; i < 10; ++i) th[i] = theta[i];
  theta_to_inertia_vector<double>(th, out);
  for (int i = 0; i < 10; ++i) inertia[i] = out[i];
  for (int t = 0; t < 10; ++t) {
    Dual<double> dth[10], dout[10];
    for (int i = 0; i < 10; ++i) dth[i] = Dual<double>(theta[i], i == t ? 1.0 : 0.0);
    theta_to_inertia_vector<Dual<double>>(dth, dout);
    double s = 0;
    for (int i = 0; i < 10; ++i) s += g_inertia[i] * dout[i].d;
    grad_theta[t] = s;
  }
  return 0;
}
int emul_elbow_loss_f64(const double* x, const double* xp, const double* inertia, const double* mu,
                        const double* half, const double* kin, double dt, double eps, int64_t B, double* loss,
                        double* force, int32_t* iters, double* grad) {
  ElbowParams<double> P;
  elbow_params_init(P, inertia, mu, half, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int i = 0; i < EL_NPARAM; ++i) if (grad) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = elbow_loss_sample(P, cfg, x + 15 * b, xp + 15 * b, (const double*)nullptr, grad,
                                force ? force + 24 * b : nullptr, (double*)nullptr, &it);
    if (iters) iters[b] = it;
  }
  return 0;
}
// the closed-form (composite body + hinge column) formulation the wavefront kernel runs (cn_elbow_wf.cuh)
int emul_elbow_loss_wf_f64(const double* x, const double* xp, const double* inertia, const double* mu,
                           const double* half, const double* kin, const double* pts, double dt, double eps, int64_t B,
                           double* loss, double* force, int32_t* iters, double* grad, double* grad_pts) {
  ElbowParams<double> P;
  double h0[6] = {0, 0, 0, 0, 0, 0};
  elbow_params_init(P, inertia, mu, half ? half : h0, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int i = 0; i < EL_NPARAM; ++i) if (grad) grad[i] = 0;
  for (int64_t b = 0; b < B; ++b) {
    int it;
    loss[b] = elbow_loss_sample_wf<double>(P, cfg, x + 15 * b, xp + 15 * b, pts ? pts + 24 * b : nullptr, grad,
                                           force ? force + 24 * b : nullptr, grad_pts ? grad_pts + 24 * b : nullptr, &it);
    if (iters) iters[b] = it;
  }
  return 0;
}
// the rollout kernels' step (closed-form mass terms, cn_elbow_wf.cuh)
int emul_elbow_step_wf_f64(const double* x, const double* inertia, const double* mu, const double* half,
                           const double* kin, const double* pts, double dt, double eps, int64_t B, double* xn,
                           double* force, double* usol) {
  ElbowParams<double> P;
  elbow_params_init<double>(P, inertia, mu, half, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int64_t b = 0; b < B; ++b)
    elbow_step_sample_wf<double>(P, cfg, x + 15 * b, pts ? pts + 24 * b : nullptr, xn + 15 * b,
                                 force ? force + 24 * b : nullptr, usol ? usol + 7 * b : nullptr);
  return 0;
}
int emul_elbow_step_f64(const double* x, const double* inertia, const double* mu, const double* half,
                        const double* kin, double dt, double eps, int64_t B, double* xn, double* force, int32_t* iters) {
  ElbowParams<double> P;
  elbow_params_init(P, inertia, mu, half, kin, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  for (int64_t b = 0; b < B; ++b) {
    int it = elbow_step_sample(P, cfg, x + 15 * b, (const double*)nullptr, xn + 15 * b, force ? force + 24 * b : nullptr);
    if (iters) iters[b] = it;
  }
  return 0;
}
// tangent rollout: per-sample gradients of sum_s xbar_s . x_s w.r.t. the 14 callable-level parameters and x0
int emul_cube_rollout_grad_f64(const double* x0, const double* inertia, const double* mu, const double* half,
                               double dt, double eps, int64_t B, int steps, const double* xbar, double* gparams,
                               double* gx0) {
  for (int64_t b = 0; b < B; ++b)
    for (int dir = 0; dir < CUBE_NTAN; ++dir) {
      const double g = cube_rollout_tangent<double>(inertia, mu, half, dt, eps, x0 + 13 * b, steps,
                                                    xbar + (int64_t)b * steps * 13, dir);
      if (dir < 14) gparams[14 * b + dir] = g; else gx0[13 * b + dir - 14] = g;
    }
  return 0;
}
// reverse-mode rollout backward (cn_cube_adjoint.cuh): forward rollout keeping the states and the QP optima, then the
// adjoint sweep; same outputs as emul_cube_rollout_grad_f64
int emul_cube_rollout_backward_f64(const double* x0, const double* inertia, const double* mu, const double* half,
                                   double dt, double eps, int64_t B, int steps, const double* xbar, double* gparams,
                                   double* gx0) {
  CubeParams<double> P;
  cube_params_init(P, inertia, mu, half, dt, eps);
  SolverCfg<double> cfg = default_cfg<double>();
  // (the plain forward solve: its optima are converged to rounding, measured: no difference with a polishing step)
  std::vector<double> traj((steps + 1) * 13), usol(steps * 6);
  for (int64_t b = 0; b < B; ++b) {
    for (int i = 0; i < 13; ++i) traj[i] = x0[13 * b + i];
    double warm[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < steps; ++s) {
      cube_step_sample<double>(P, cfg, &traj[13 * s], &traj[13 * (s + 1)], (double*)nullptr, warm);
      for (int i = 0; i < 6; ++i) usol[6 * s + i] = warm[i];
    }
    cube_rollout_backward_sample<double>(P, traj.data(), usol.data(), xbar + (int64_t)b * steps * 13, steps,
                                         gparams + 14 * b, gx0 + 13 * b);
  }
  return 0;
}
// the same with every step's optimum kept by a plain forward rollout (one evaluation per dual-number step)
int emul_elbow_rollout_grad_saved_f64(const double* x0, const double* inertia, const double* mu, const double* half,
                                      const double* kin, double dt, double eps, int64_t B, int steps, const double* xbar,
                                      double* gparams, double* gx0) {
  ElbowParams<double> P;
  elbow_params_init<double>(P, inertia, mu, half, kin, dt, eps);
  const SolverCfg<double> cfg = default_cfg<double>();
  std::vector<double> usol((size_t)steps * 7);
  for (int64_t b = 0; b < B; ++b) {
    double x[15], xn[15];
    for (int i = 0; i < 15; ++i) x[i] = x0[15 * b + i];
    for (int s = 0; s < steps; ++s) {
      elbow_step_sample<double>(P, cfg, x, (const double*)nullptr, xn, (double*)nullptr, &usol[(size_t)s * 7]);
      for (int i = 0; i < 15; ++i) x[i] = xn[i];
    }
    for (int dir = 0; dir < ELBOW_NTAN; ++dir) {
      const double g = elbow_rollout_tangent<double>(inertia, mu, half, kin, dt, eps, x0 + 15 * b, steps,
                                                     xbar + (int64_t)b * steps * 15, dir, usol.data());
      if (dir < ELBOW_NPARAM_TAN) gparams[ELBOW_NPARAM_TAN * b + dir] = g; else gx0[15 * b + dir - ELBOW_NPARAM_TAN] = g;
    }
  }
  return 0;
}
int emul_elbow_rollout_grad_f64(const double* x0, const double* inertia, const double* mu, const double* half,
                                const double* kin, double dt, double eps, int64_t B, int steps, const double* xbar,
                                double* gparams, double* gx0) {
  for (int64_t b = 0; b < B; ++b)
    for (int dir = 0; dir < ELBOW_NTAN; ++dir) {
      const double g = elbow_rollout_tangent<double>(inertia, mu, half, kin, dt, eps, x0 + 15 * b, steps,
                                                     xbar + (int64_t)b * steps * 15, dir);
      if (dir < ELBOW_NPARAM_TAN) gparams[ELBOW_NPARAM_TAN * b + dir] = g; else gx0[15 * b + dir - ELBOW_NPARAM_TAN] = g;
    }
  return 0;
}
}
