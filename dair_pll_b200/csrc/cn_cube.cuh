// Single floating rigid body with one box geometry against the ground plane
// (assets/contactnets_cube.urdf): n_q = 7, n_v = 6, 4 contacts (top-4 of the 8 box
// corners, geometry.py:162-202, 487-491), k = 12.
//
// Replaces, per sample, what the reference computes with ~5.5k ATen calls:
//   LagrangianTerms.forward        multibody_terms.py:214-237   (M, M^-1 F)
//   ContactTerms.forward           multibody_terms.py:428-521   (phi, J)
//   collide_plane_convex / top-k   geometry.py:553-582, 162-202
//   contactnets_loss               multibody_learnable_system.py:104-197
//   forward_dynamics               multibody_learnable_system.py:199-304
//   SAPSolver.apply                (un-vendored; restated, see oracle/cone_qp.c)
//
// Internal velocity coordinates are the WORLD-frame twist u^ = [R w_B ; v_W]: in
// them a contact Jacobian is [-S(rho_c), I3] with rho_c the world-frame lever arm,
// so J u, J^T f and J^T K J reduce to cross products and no dense 12x6 J is ever
// formed; the mass matrix becomes [[R I_o R^T, m S(R c)], [-m S(R c), m I3]] and its
// inverse is closed-form (Schur complement = R I_sym R^T).  The QP is solved in
// primal (velocity) form:  min_u 1/2 u^T M u + eps/2 sum_c |Pi(-(D_mu J_c u + q_c)/eps)|^2
// by Newton with a derivative-based line search; f_c = Pi(...) at the optimum.
#pragma once
#include "cn_common.cuh"

#ifndef CN_STAT_LS
#define CN_STAT_LS()
#define CN_STAT_UNIT()
#endif

namespace cn {

constexpr int CUBE_NQ = 7, CUBE_NV = 6, CUBE_NX = 13, CUBE_NC = 4, CUBE_K = 12;
constexpr int CUBE_NPARAM = 14;  // [inertia 10 | mu_pair 1 | half_lengths 3]

template <typename T> struct CubeParams {
  T m, c[3], Isym[6];   // the 10-vector the generated callables take (multibody_terms.py:198-205)
  T mu;                 // combined pair friction 2 mu_a mu_b/(mu_a+mu_b) (multibody_terms.py:471)
  T h[3];               // |length_params| (geometry.py:394-397)
  T dt, eps, inv_eps, grav;
  // derived
  T Io[6];              // I_sym - m S(c)^2  (inertia about the body origin)
  T Isym_inv[6];
  T inv_m;
  T dscale[6];          // 1/diag(M) in body coordinates, for the scaled stopping test
};

template <typename T>
CN_HD void cube_params_init(CubeParams<T>& P, const T* inertia, const T* mu, const T* half, T dt, T eps) {
  P.m = inertia[0];
  for (int i = 0; i < 3; ++i) P.c[i] = inertia[1 + i];
  for (int i = 0; i < 6; ++i) P.Isym[i] = inertia[4 + i];
  P.mu = mu[0];
  for (int i = 0; i < 3; ++i) P.h[i] = half[i];
  P.dt = dt; P.eps = eps; P.inv_eps = T(1) / eps; P.grav = T(9.81);
  const T cx = P.c[0], cy = P.c[1], cz = P.c[2], m = P.m;
  // -S(c)^2 = |c|^2 I - c c^T
  P.Io[0] = P.Isym[0] + m * (cy * cy + cz * cz);
  P.Io[1] = P.Isym[1] + m * (cx * cx + cz * cz);
  P.Io[2] = P.Isym[2] + m * (cx * cx + cy * cy);
  P.Io[3] = P.Isym[3] - m * cx * cy;
  P.Io[4] = P.Isym[4] - m * cx * cz;
  P.Io[5] = P.Isym[5] - m * cy * cz;
  sym3_inv(P.Isym, P.Isym_inv);
  P.inv_m = T(1) / m;
  P.dscale[0] = T(1) / P.Io[0]; P.dscale[1] = T(1) / P.Io[1]; P.dscale[2] = T(1) / P.Io[2];
  P.dscale[3] = P.dscale[4] = P.dscale[5] = P.inv_m;
}

// Per-sample quantities that stay fixed during the Newton solve.
template <typename T> struct CubeProblem {
  T R[9];        // world <- body rotation
  T IW[6];       // R Io R^T
  T mcW[3];      // m R c
  T rho[12];     // world-frame lever arms of the 4 selected corners
  T q[12];       // QP linear term, sappy order [tx, ty, n] per contact
  uint32_t sel;  // 3 sign bits per selected corner (bit set = +h), corner c at bits 3c..3c+2 (x,y,z)
};

// apply the body-frame closed-form inverse mass matrix:  [aw; av] = M^-1 [tau; frc]
// (tau, aw body frame; frc, av world frame).  Schur complement of M is I_sym.
template <typename T>
CN_HD void cube_minv(const CubeParams<T>& P, const T* R, const T* tau, const T* frc, T* aw, T* av) {
  T fb[3], cxf[3], rhs[3], cxa[3], rc[3];
  rot3t(R, frc, fb);
  cross3(P.c, fb, cxf);
  for (int i = 0; i < 3; ++i) rhs[i] = tau[i] - cxf[i];
  sym3_mul(P.Isym_inv, rhs, aw);
  cross3(P.c, aw, cxa);
  rot3(R, cxa, rc);
  for (int i = 0; i < 3; ++i) av[i] = frc[i] * P.inv_m + rc[i];
}

// u^ -> M^ u^ in world-twist coordinates
template <typename T>
CN_HD void cube_mass_mul(const CubeParams<T>& P, const CubeProblem<T>& S, const T* u, T* o) {
  T a[3], b[3], c[3];
  sym3_mul(S.IW, u, a);
  cross3(S.mcW, u + 3, b);
  cross3(S.mcW, u, c);
  for (int i = 0; i < 3; ++i) { o[i] = a[i] + b[i]; o[3 + i] = P.m * u[3 + i] - c[i]; }
}

// Select the 4 corners with the largest support in direction d (body frame) -- top-k of
// d . (sigma o h) over the 8 sign patterns (geometry.py:191-197).  Returned in ascending
// vertex-index order (the reference's order is unspecified: topk(sorted=False)).
template <typename T> CN_HD uint32_t cube_select_corners(const T* d, const T* h) {
  const T a0 = d[0] * h[0], a1 = d[1] * h[1], a2 = d[2] * h[2];
  T dots[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    dots[i] = ((i & 4) ? a0 : -a0) + ((i & 2) ? a1 : -a1) + ((i & 1) ? a2 : -a2);
  uint32_t sel = 0; int n = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) rank += (dots[j] > dots[i]) || (dots[j] == dots[i] && j < i);
    if (rank < 4) {
      // vertex i = (x: bit2, y: bit1, z: bit0) (geometry.py:39-41)
      const uint32_t bits = ((i >> 2) & 1) | (((i >> 1) & 1) << 1) | ((i & 1) << 2);
      sel |= bits << (3 * n);
      ++n;
    }
  }
  return sel;
}

template <typename T> CN_HD T sgn_bit(uint32_t sel, int c, int axis) {
  return ((sel >> (3 * c + axis)) & 1u) ? T(1) : T(-1);
}

// Shared geometry set-up: R, world inertia, corner selection, lever arms, phi.
template <typename T>
CN_HD void cube_geometry(const CubeParams<T>& P, const T* quat, T pos_z, CubeProblem<T>& S, T* phi) {
  quat_to_rot(quat, S.R);
  const T* R = S.R;
  // IW = R Io R^T
  T A[9];
  for (int i = 0; i < 3; ++i) {
    const T r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
    A[3 * i + 0] = r0 * P.Io[0] + r1 * P.Io[3] + r2 * P.Io[4];
    A[3 * i + 1] = r0 * P.Io[3] + r1 * P.Io[1] + r2 * P.Io[5];
    A[3 * i + 2] = r0 * P.Io[4] + r1 * P.Io[5] + r2 * P.Io[2];
  }
  S.IW[0] = A[0] * R[0] + A[1] * R[1] + A[2] * R[2];
  S.IW[1] = A[3] * R[3] + A[4] * R[4] + A[5] * R[5];
  S.IW[2] = A[6] * R[6] + A[7] * R[7] + A[8] * R[8];
  S.IW[3] = A[0] * R[3] + A[1] * R[4] + A[2] * R[5];
  S.IW[4] = A[0] * R[6] + A[1] * R[7] + A[2] * R[8];
  S.IW[5] = A[3] * R[6] + A[4] * R[7] + A[5] * R[8];
  T cW[3];
  rot3(R, P.c, cW);
  for (int i = 0; i < 3; ++i) S.mcW[i] = P.m * cW[i];
  // support direction in the body frame: -(third row of R)  (geometry.py:560-564)
  const T d[3] = {-R[6], -R[7], -R[8]};
  S.sel = cube_select_corners(d, P.h);
  T Rh[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) Rh[3 * i + k] = R[3 * i + k] * P.h[k];
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T sx = sgn_bit<T>(S.sel, c, 0), sy = sgn_bit<T>(S.sel, c, 1), sz = sgn_bit<T>(S.sel, c, 2);
    for (int i = 0; i < 3; ++i) S.rho[3 * c + i] = sx * Rh[3 * i] + sy * Rh[3 * i + 1] + sz * Rh[3 * i + 2];
    phi[c] = S.rho[3 * c + 2] + pos_z;   // geometry.py:568-571
  }
}

// Contact-free acceleration M^-1 F at (R, w_B, ...), in body(w)/world(v) state coordinates
// (closed form of the generated lagrangian_forces, see oracle/callables.py).
template <typename T>
CN_HD void cube_free_accel(const CubeParams<T>& P, const T* R, const T* wB, T* aw, T* av) {
  T Iw[3], wIw[3], gB[3], cg[3], tau[3], wc[3], wwc[3], Rwwc[3], frc[3];
  sym3_mul(P.Io, wB, Iw);
  cross3(wB, Iw, wIw);
  gB[0] = -P.grav * R[6]; gB[1] = -P.grav * R[7]; gB[2] = -P.grav * R[8];
  cross3(P.c, gB, cg);
  for (int i = 0; i < 3; ++i) tau[i] = -wIw[i] + P.m * cg[i];
  cross3(wB, P.c, wc);
  cross3(wB, wc, wwc);
  rot3(R, wwc, Rwwc);
  for (int i = 0; i < 3; ++i) frc[i] = -P.m * Rwwc[i];
  frc[2] -= P.m * P.grav;
  cube_minv(P, R, tau, frc, aw, av);
}

// Gradient (and optionally Hessian) of the primal objective at u (world twist).
//   g = M u - sum_c J_c^T f~_c ;  H = M + sum_c J_c^T K_c J_c
// Also returns the scaled norms used by the stopping test and, if f_out, the forces.
template <typename T, bool WANT_H>
CN_HD void cube_eval(const CubeParams<T>& P, const CubeProblem<T>& S, const T* u, T* g, T* H,
                     T* f_out, T& res2, T& scale2) {
  T Mu[6];
  cube_mass_mul(P, S, u, Mu);
  T z[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  if (WANT_H) {
    // H starts as M^
    H[0] = S.IW[0]; H[7] = S.IW[1]; H[14] = S.IW[2];
    H[6] = S.IW[3]; H[12] = S.IW[4]; H[13] = S.IW[5];
    // M_vw = -S(mcW): rows 3..5, cols 0..2
    H[18] = T(0);       H[19] = S.mcW[2];  H[20] = -S.mcW[1];
    H[24] = -S.mcW[2];  H[25] = T(0);      H[26] = S.mcW[0];
    H[30] = S.mcW[1];   H[31] = -S.mcW[0]; H[32] = T(0);
    H[21] = P.m; H[28] = P.m; H[35] = P.m; H[27] = T(0); H[33] = T(0); H[34] = T(0);
  }
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T* rho = S.rho + 3 * c;
    T e[3], r[3], f[3], K[6];
    cross3(u, rho, e);
    r[0] = P.mu * (e[0] + u[3]) + S.q[3 * c];
    r[1] = P.mu * (e[1] + u[4]) + S.q[3 * c + 1];
    r[2] = (e[2] + u[5]) + S.q[3 * c + 2];
    cone_eval<T, WANT_H>(r, P.inv_eps, P.mu, f, K);
    if (f_out) { f_out[3 * c] = f[0]; f_out[3 * c + 1] = f[1]; f_out[3 * c + 2] = f[2]; }
    const T ft[3] = {P.mu * f[0], P.mu * f[1], f[2]};
    T tq[3];
    cross3(rho, ft, tq);
    for (int i = 0; i < 3; ++i) { z[i] += tq[i]; z[3 + i] += ft[i]; }
    if (WANT_H) {
      // Kfull rows
      const T K0[3] = {K[0], K[1], K[2]}, K1[3] = {K[1], K[3], K[4]}, K2[3] = {K[2], K[4], K[5]};
      // P = S(rho) K : column j = rho x K_j  (K symmetric)
      T P0[3], P1[3], P2[3];
      cross3(rho, K0, P0); cross3(rho, K1, P1); cross3(rho, K2, P2);
      // H_vw (rows 3+j, cols i) += P^T  i.e. H[(3+j)*6 + i] += P_j[i]
      for (int i = 0; i < 3; ++i) { H[18 + i] += P0[i]; H[24 + i] += P1[i]; H[30 + i] += P2[i]; }
      // H_ww += -P S(rho): row i = rho x (row i of P); row i of P = (P0[i], P1[i], P2[i])
      const T r0[3] = {P0[0], P1[0], P2[0]}, r1[3] = {P0[1], P1[1], P2[1]}, r2[3] = {P0[2], P1[2], P2[2]};
      T w0[3], w1[3], w2[3];
      cross3(rho, r0, w0); cross3(rho, r1, w1); cross3(rho, r2, w2);
      H[0] += w0[0]; H[6] += w1[0]; H[7] += w1[1]; H[12] += w2[0]; H[13] += w2[1]; H[14] += w2[2];
      // H_vv += K
      H[21] += K[0]; H[27] += K[1]; H[28] += K[3]; H[33] += K[2]; H[34] += K[4]; H[35] += K[5];
    }
  }
  // scaled residual test quantities: D = 1/diag(M) (body-frame diagonal as a fixed scale)
  res2 = T(0); T a2 = T(0), b2 = T(0);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    g[i] = Mu[i] - z[i];
    res2 += g[i] * g[i] * P.dscale[i];
    a2 += Mu[i] * Mu[i] * P.dscale[i];
    b2 += z[i] * z[i] * P.dscale[i];
  }
  scale2 = t_max(a2, b2);
}

// 1-D derivative phi'(alpha) (and curvature) along u + alpha d, given the residuals r1 at
// alpha = 1 and the direction images ed_c = D_mu J_c d:  r(alpha) = r1 - (1-alpha) ed.
template <typename T>
CN_HD void cube_line(const CubeParams<T>& P, const T* r1, const T* ed, T uMd, T dMd, T alpha, T& d1, T& d2) {
  d1 = uMd + alpha * dMd;
  d2 = dMd;
  const T back = T(1) - alpha;
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    T r[3], f[3], K[6];
    for (int j = 0; j < 3; ++j) r[j] = r1[3 * c + j] - back * ed[3 * c + j];
    // cone_eval's K carries D_mu; the line needs the plain G/eps on ed (already D_mu-scaled),
    // so evaluate with mu = 1.
    cone_eval<T, true>(r, P.inv_eps, T(1), f, K);
    const T* e = ed + 3 * c;
    d1 -= e[0] * f[0] + e[1] * f[1] + e[2] * f[2];
    d2 += e[0] * (K[0] * e[0] + K[1] * e[1] + K[2] * e[2]) + e[1] * (K[1] * e[0] + K[3] * e[1] + K[4] * e[2]) +
          e[2] * (K[2] * e[0] + K[4] * e[1] + K[5] * e[2]);
  }
}

// True iff u = 0 is already optimal: every contact's y = -q/eps lies in the polar cone, so
// all forces vanish and the gradient M*0 - J^T 0 is exactly zero (free flight).
template <typename T> CN_HD bool cube_trivially_solved(const CubeProblem<T>& S) {
  bool open = true;
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T qt2 = S.q[3 * c] * S.q[3 * c] + S.q[3 * c + 1] * S.q[3 * c + 1];
    const T qn = S.q[3 * c + 2];
    open = open && (qn >= T(0)) && (qt2 <= qn * qn);
  }
  return open;
}

template <typename T> CN_HD bool cube_converged(const SolverCfg<T>& cfg, T res2, T scale2) {
  return !(res2 > cfg.tol_rel * cfg.tol_rel * scale2);   // also true for NaN
}

// One schedulable unit of the Newton solve (the wavefront kernel runs one unit per lane per
// trip): Hessian at u, Cholesky step, trial point, derivative line search if the full step
// overshoots.  Updates u / prev_res2 / it; returns true when the sample is finished.
template <typename T>
CN_HD bool cube_newton_unit(const CubeParams<T>& P, const CubeProblem<T>& S, const SolverCfg<T>& cfg, T* u,
                            T& prev_res2, int& it) {
  T g[6], H[36], res2, scale2;
  CN_STAT_UNIT();
  cube_eval<T, true>(P, S, u, g, H, (T*)nullptr, res2, scale2);
  if (cube_converged(cfg, res2, scale2)) return true;
  if (res2 <= cfg.tol_stall * cfg.tol_stall * scale2 && res2 >= T(0.25) * prev_res2) return true;  // rounding floor
  if (it >= cfg.max_iter) return true;
  prev_res2 = res2;
  T d[6];
  chol_solve_neg<T, 6>(H, g, d);
  T d0 = T(0);
#pragma unroll
  for (int i = 0; i < 6; ++i) d0 += g[i] * d[i];
  T u1[6], g1[6], res2b, scale2b;
#pragma unroll
  for (int i = 0; i < 6; ++i) u1[i] = u[i] + d[i];
  cube_eval<T, false>(P, S, u1, g1, (T*)nullptr, (T*)nullptr, res2b, scale2b);
  T d1 = T(0);
#pragma unroll
  for (int i = 0; i < 6; ++i) d1 += g1[i] * d[i];
  const T thresh = cfg.ls_c * t_abs(d0);
  ++it;
  if (d1 <= thresh) {
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = u1[i];
    return cube_converged(cfg, res2b, scale2b);
  }
  // overshoot: safeguarded Newton on phi' over (0,1)
  T Md[6], uMd = T(0), dMd = T(0);
  cube_mass_mul(P, S, d, Md);
#pragma unroll
  for (int i = 0; i < 6; ++i) { uMd += u[i] * Md[i]; dMd += d[i] * Md[i]; }
  T r1[12], ed[12];
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    T e[3], e1[3];
    cross3(d, S.rho + 3 * c, e);
    cross3(u1, S.rho + 3 * c, e1);
    ed[3 * c] = P.mu * (e[0] + d[3]); ed[3 * c + 1] = P.mu * (e[1] + d[4]); ed[3 * c + 2] = e[2] + d[5];
    r1[3 * c] = P.mu * (e1[0] + u1[3]) + S.q[3 * c];
    r1[3 * c + 1] = P.mu * (e1[1] + u1[4]) + S.q[3 * c + 1];
    r1[3 * c + 2] = (e1[2] + u1[5]) + S.q[3 * c + 2];
  }
  T lo = T(0), hi = T(1), alpha, da, ha;
  cube_line(P, r1, ed, uMd, dMd, T(1), da, ha);       // first guess: Newton step on phi' from alpha = 1
  alpha = T(1) - da / ha;
  if (!(alpha > lo && alpha < hi)) alpha = T(0.5);
  for (int ls = 0; ls < 40; ++ls) {
    cube_line(P, r1, ed, uMd, dMd, alpha, da, ha);
    CN_STAT_LS();
    if (t_abs(da) <= thresh) break;
    if (da < T(0)) lo = alpha; else hi = alpha;
    T an = alpha - da / ha;
    if (!(an > lo && an < hi)) an = T(0.5) * (lo + hi);
    if (hi - lo <= T(4) * eps_of<T>() * hi) { alpha = lo > T(0) ? lo : an; break; }
    alpha = an;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) u[i] += alpha * d[i];
  return false;
}

// Forces f = Pi(-(D_mu J u + q)/eps) at u (sappy order).
template <typename T>
CN_HD void cube_forces(const CubeParams<T>& P, const CubeProblem<T>& S, const T* u, T* f) {
  T g[6], res2, scale2;
  cube_eval<T, false>(P, S, u, g, (T*)nullptr, f, res2, scale2);
}

// Newton solve from u (in: start point, out: optimum).  Returns iterations; f = forces at u.
template <typename T>
CN_HD int cube_solve(const CubeParams<T>& P, const CubeProblem<T>& S, const SolverCfg<T>& cfg, T* u, T* f) {
  int it = 0;
  if (!cube_trivially_solved(S)) {
    T prev = T(-1);
    while (!cube_newton_unit(P, S, cfg, u, prev, it)) {}
  }
  cube_forces(P, S, u, f);
  return it;
}

// ---------------------------------------------------------------------------
// ContactNets loss (multibody_learnable_system.py:104-197): prologue builds the QP
// at (q+, v+); epilogue evaluates the loss at the solved (detached) forces and, by
// the envelope theorem (:172-175), the parameter gradient in the same pass.
// ---------------------------------------------------------------------------
template <typename T> struct CubeLossAux {
  T dv[6];      // v+ - (v + dt a), state coordinates [body w ; world v]   (:156)
  T acc[6];     // contact-free acceleration, state coordinates
  T vp[6];      // v+, state coordinates
  T phi[4];
  T konst;      // 1/2 dv^T M dv + sum max(-phi,0)^2   (:163-170)
};

template <typename T>
CN_HD void cube_loss_prologue(const CubeParams<T>& P, const T* x, const T* xp, CubeProblem<T>& S,
                              CubeLossAux<T>& A) {
  cube_geometry(P, xp, xp[6], S, A.phi);
  for (int i = 0; i < 6; ++i) A.vp[i] = xp[7 + i];
  cube_free_accel(P, S.R, A.vp, A.acc, A.acc + 3);
  for (int i = 0; i < 6; ++i) A.dv[i] = A.vp[i] - (x[7 + i] + P.dt * A.acc[i]);
  // world twists
  T dvW[6], vW[6];
  rot3(S.R, A.dv, dvW); rot3(S.R, A.vp, vW);
  for (int i = 0; i < 3; ++i) { dvW[3 + i] = A.dv[3 + i]; vW[3 + i] = A.vp[3 + i]; }
  T pen = T(0);
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T* rho = S.rho + 3 * c;
    T ed[3], ev[3];
    cross3(dvW, rho, ed); cross3(vW, rho, ev);
    for (int i = 0; i < 3; ++i) { ed[i] += dvW[3 + i]; ev[i] += vW[3 + i]; }
    const T sx = P.mu * ev[0], sy = P.mu * ev[1];
    const T speed = t_sqrt(sx * sx + sy * sy);
    S.q[3 * c] = -P.mu * ed[0] + P.dt * sx;                       // :158-161
    S.q[3 * c + 1] = -P.mu * ed[1] + P.dt * sy;
    S.q[3 * c + 2] = -ed[2] + t_abs(A.phi[c]) + P.dt * speed;
    const T pneg = t_max(-A.phi[c], T(0));
    pen += pneg * pneg;
  }
  T Mdv[6];
  cube_mass_mul(P, S, dvW, Mdv);
  T e = T(0);
  for (int i = 0; i < 6; ++i) e += dvW[i] * Mdv[i];
  A.konst = T(0.5) * e + pen;
}

// vex(X) with <X, S(a)> = a . vex(X);  X row-major 3x3
template <typename T> CN_HD void vex3(const T* X, T* o) {
  o[0] = X[7] - X[5]; o[1] = X[2] - X[6]; o[2] = X[3] - X[1];
}

// Loss value and (if grad != nullptr) += d loss / d [inertia(10), mu, half(3)].
// f: solved forces, sappy order.  force_out (nullable): reference order [n(4); (tx,ty)(4)].
template <typename T>
CN_HD T cube_loss_epilogue(const CubeParams<T>& P, const CubeProblem<T>& S, const CubeLossAux<T>& A,
                           const T* f_in, T* grad, T* force_out) {
  T f[12];
  bool bad = false;
  for (int i = 0; i < 12; ++i) {
    f[i] = f_in[i];
    bad = bad || !(t_abs(f[i]) <= T(1e3));     // also catches NaN / Inf   (:186-189)
  }
  if (bad) {
    if (force_out) for (int i = 0; i < 12; ++i) force_out[i] = T(0);
    return T(0);                               // force := 0, constant := 0 (:191-192) => loss 0, grad 0
  }
  if (force_out) {
    for (int c = 0; c < 4; ++c) {
      force_out[c] = f[3 * c + 2];
      force_out[4 + 2 * c] = f[3 * c];
      force_out[4 + 2 * c + 1] = f[3 * c + 1];
    }
  }
  const T* R = S.R;
  // z = J^T f in world twist, then state coordinates
  T zW[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  T qf = T(0), ff = T(0);
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T ft[3] = {P.mu * f[3 * c], P.mu * f[3 * c + 1], f[3 * c + 2]};
    T tq[3];
    cross3(S.rho + 3 * c, ft, tq);
    for (int i = 0; i < 3; ++i) { zW[i] += tq[i]; zW[3 + i] += ft[i]; }
    for (int i = 0; i < 3; ++i) { qf += S.q[3 * c + i] * f[3 * c + i]; ff += f[3 * c + i] * f[3 * c + i]; }
  }
  T z[6], y[6];
  rot3t(R, zW, z);
  for (int i = 0; i < 3; ++i) z[3 + i] = zW[3 + i];
  cube_minv(P, R, z, z + 3, y, y + 3);
  T zy = T(0);
  for (int i = 0; i < 6; ++i) zy += z[i] * y[i];
  const T loss = T(0.5) * zy + T(0.5) * P.eps * ff + qf + A.konst;   // :194-195
  if (!grad) return loss;

  // ---- envelope backward (forces fixed) ----
  const T* dv = A.dv; const T* a = A.acc; const T* w = A.vp;   // w = body angular velocity (first 3)
  T lam[6], b[6];
  for (int i = 0; i < 6; ++i) { b[i] = y[i] - dv[i]; lam[i] = P.dt * b[i]; }   // lam = -dt (dv - y)
  // Mbar = -1/2 y y^T + 1/2 dv dv^T - lam a^T
  T Kww[9], N[9], trvv = T(0);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Kww[3 * i + j] = T(0.5) * (dv[i] * dv[j] - y[i] * y[j]) - lam[i] * a[j];
      N[3 * i + j] = (dv[i] * dv[3 + j] - y[i] * y[3 + j]) - lam[i] * a[3 + j] - lam[3 + j] * a[i];
    }
  for (int i = 0; i < 3; ++i) trvv += T(0.5) * (dv[3 + i] * dv[3 + i] - y[3 + i] * y[3 + i]) - lam[3 + i] * a[3 + i];
  // F-term into the I_o adjoint: (w x lam_w) w^T
  T wxl[3];
  cross3(w, lam, wxl);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Kww[3 * i + j] += wxl[i] * w[j];
  // I_sym
  grad[4] += Kww[0]; grad[5] += Kww[4]; grad[6] += Kww[8];
  grad[7] += Kww[1] + Kww[3]; grad[8] += Kww[2] + Kww[6]; grad[9] += Kww[5] + Kww[7];
  // helpers
  const T* c = P.c;
  const T trK = Kww[0] + Kww[4] + Kww[8];
  T Kc[3], Ktc[3];
  for (int i = 0; i < 3; ++i) {
    Kc[i] = Kww[3 * i] * c[0] + Kww[3 * i + 1] * c[1] + Kww[3 * i + 2] * c[2];
    Ktc[i] = Kww[i] * c[0] + Kww[3 + i] * c[1] + Kww[6 + i] * c[2];
  }
  const T cc = dot3(c, c);
  T NR[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) NR[3 * i + j] = N[3 * i] * R[j] + N[3 * i + 1] * R[3 + j] + N[3 * i + 2] * R[6 + j];
  T vNR[3];
  vex3(NR, vNR);
  T gB[3] = {-P.grav * R[6], -P.grav * R[7], -P.grav * R[8]};
  T ell[3];
  rot3t(R, lam + 3, ell);
  T cxg[3], wc[3], wwc[3], wl[3], wwl[3], gxl[3];
  cross3(c, gB, cxg);
  cross3(w, c, wc); cross3(w, wc, wwc);
  cross3(w, ell, wl); cross3(w, wl, wwl);
  cross3(gB, lam, gxl);
  // mass
  grad[0] += cc * trK - dot3(c, Kc) + dot3(c, vNR) + trvv + dot3(lam, cxg) - dot3(ell, wwc) - P.grav * lam[5];
  // com
  for (int i = 0; i < 3; ++i)
    grad[1 + i] += P.m * (T(2) * trK * c[i] - Kc[i] - Ktc[i] + vNR[i] + gxl[i] - wwl[i]);
  // contacts: mu and half lengths
  T bW[3], wW[3];
  rot3(R, b, bW); rot3(R, w, wW);
  T gmu = T(0), gh[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int cidx = 0; cidx < CUBE_NC; ++cidx) {
    const T* rho = S.rho + 3 * cidx;
    T eb[3], ev[3];
    cross3(bW, rho, eb); cross3(wW, rho, ev);
    for (int i = 0; i < 3; ++i) { eb[i] += b[3 + i]; ev[i] += A.vp[3 + i]; }
    const T ftx = f[3 * cidx], fty = f[3 * cidx + 1], fn = f[3 * cidx + 2];
    const T sx = P.mu * ev[0], sy = P.mu * ev[1];
    const T speed = t_sqrt(sx * sx + sy * sy);
    const T ux = speed > T(0) ? sx / speed : T(0), uy = speed > T(0) ? sy / speed : T(0);
    const T gx = P.dt * (fn * ux + ftx), gy = P.dt * (fn * uy + fty);
    gmu += ftx * eb[0] + fty * eb[1] + gx * ev[0] + gy * ev[1];
    const T ft[3] = {P.mu * ftx, P.mu * fty, fn};
    const T gt[3] = {P.mu * gx, P.mu * gy, T(0)};
    T ftB[3], gtB[3], p1[3], p2[3];
    rot3t(R, ft, ftB); rot3t(R, gt, gtB);
    cross3(ftB, b, p1); cross3(gtB, w, p2);
    const T phic = A.phi[cidx];
    const T phibar = (phic > T(0) ? fn : (phic < T(0) ? -fn : T(0))) - T(2) * t_max(-phic, T(0));
    for (int k = 0; k < 3; ++k) gh[k] += sgn_bit<T>(S.sel, cidx, k) * (p1[k] + p2[k] + phibar * R[6 + k]);
  }
  grad[10] += gmu;
  for (int k = 0; k < 3; ++k) grad[11 + k] += gh[k];
  return loss;
}

// Whole per-sample loss path (the simple, one-thread-per-sample composition).
template <typename T>
CN_HD T cube_loss_sample(const CubeParams<T>& P, const SolverCfg<T>& cfg, const T* x, const T* xp, T* grad,
                         T* force_out, int* iters_out) {
  CubeProblem<T> S;
  CubeLossAux<T> A;
  cube_loss_prologue(P, x, xp, S, A);
  T u[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, f[12];
  const int it = cube_solve(P, S, cfg, u, f);
  if (iters_out) *iters_out = it;
  return cube_loss_epilogue(P, S, A, f, grad, force_out);
}

// ---------------------------------------------------------------------------
// Learnable time step: forward_dynamics (multibody_learnable_system.py:260-304) +
// VelocityIntegrator.step (integrator.py:153-162) + FloatingBaseSpace.exponential
// (state_space.py:466-486, quaternion.py:89-104, 276-309).  No quaternion renormalisation.
// ---------------------------------------------------------------------------
template <typename T>
CN_HD int cube_step_sample(const CubeParams<T>& P, const SolverCfg<T>& cfg, const T* x, T* xn, T* force_out) {
  CubeProblem<T> S;
  T phi[4], acc[6], vm[6], vmW[6];
  cube_geometry(P, x, x[6], S, phi);
  cube_free_accel(P, S.R, x + 7, acc, acc + 3);
  for (int i = 0; i < 6; ++i) vm[i] = x[7 + i] + P.dt * acc[i];            // :285
  rot3(S.R, vm, vmW);
  for (int i = 0; i < 3; ++i) vmW[3 + i] = vm[3 + i];
  const T inv_dt = T(1) / P.dt;
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    T e[3];
    cross3(vmW, S.rho + 3 * c, e);
    S.q[3 * c] = P.mu * (e[0] + vmW[3]);                                   // :286
    S.q[3 * c + 1] = P.mu * (e[1] + vmW[4]);
    S.q[3 * c + 2] = (e[2] + vmW[5]) + phi[c] * inv_dt;
  }
  T u[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, f[12];
  const int it = cube_solve(P, S, cfg, u, f);
  if (force_out)
    for (int c = 0; c < 4; ++c) {
      force_out[c] = f[3 * c + 2]; force_out[4 + 2 * c] = f[3 * c]; force_out[4 + 2 * c + 1] = f[3 * c + 1];
    }
  // v+ = v- + M^-1 J^T f = v- + u   (u is exactly M^-1 J^T f at the optimum; use the
  // force form so the result is a function of f as in the reference, :303-304)
  T zW[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    const T ft[3] = {P.mu * f[3 * c], P.mu * f[3 * c + 1], f[3 * c + 2]};
    T tq[3];
    cross3(S.rho + 3 * c, ft, tq);
    for (int i = 0; i < 3; ++i) { zW[i] += tq[i]; zW[3 + i] += ft[i]; }
  }
  T z[6], y[6], vn[6];
  rot3t(S.R, zW, z);
  for (int i = 0; i < 3; ++i) z[3 + i] = zW[3 + i];
  cube_minv(P, S.R, z, z + 3, y, y + 3);
  for (int i = 0; i < 6; ++i) vn[i] = vm[i] + y[i];
  // q+ = q (+) v+ dt
  const T rx = vn[0] * P.dt, ry = vn[1] * P.dt, rz = vn[2] * P.dt;
  const T ang = t_sqrt(rx * rx + ry * ry + rz * rz);
  const T half = T(0.5) * ang;
  const T sinc = half > T(0) ? sin(half) / half : T(1);      // quaternion.py:208-229
  const T dw = cos(half), k = T(0.5) * sinc;
  const T dx = rx * k, dy = ry * k, dz = rz * k;
  const T qw = x[0], qx = x[1], qy = x[2], qz = x[3];
  xn[0] = qw * dw - (qx * dx + qy * dy + qz * dz);
  xn[1] = qw * dx + dw * qx + (qy * dz - qz * dy);
  xn[2] = qw * dy + dw * qy + (qz * dx - qx * dz);
  xn[3] = qw * dz + dw * qz + (qx * dy - qy * dx);
  for (int i = 0; i < 3; ++i) xn[4 + i] = x[4 + i] + vn[3 + i] * P.dt;
  for (int i = 0; i < 6; ++i) xn[7 + i] = vn[i];
  return it;
}

}  // namespace cn
