// Reverse-mode derivative of the learnable time step of the cube: what the reference
// obtains by autograd through forward_dynamics and sappy's backward (multibody_learnable_system.py:293-304;
// integrator.py:153-162; state_space.py:466-486), hand-derived for the step code of cn_cube.cuh.
//
// One step maps (x, theta) -> x+ through
//   R(quat);  I_W = R Io R^T, mc_W = m R c, rho_c = R (sigma_c o h)            (cube_geometry)
//   a = M^-1 F(R, w)                                                           (cube_free_accel)
//   vm = v + dt a;  vm_W = [R vm_w ; vm_v]
//   q_c = D_mu (vm_W,w x rho_c + vm_W,v) + [0, 0, (rho_c,z + pos_z)/dt]         (cube_step_prologue)
//   u* = argmin_u  1/2 u^T M^ u + eps/2 sum_c |Pi(-(D_mu J_c u + q_c)/eps)|^2   (the cone QP, primal form)
//   vn = [vm_w + R^T u_w ; vm_v + u_v];  quat+ = quat (x) exp(vn_w dt), pos+ = pos + vn_v dt
// Given the cotangent of x+ the adjoint runs these lines backwards.  The QP is differentiated implicitly: with
// g(u; data) = M^ u - sum_c J_c^T D_mu f_c = 0 at u*, lambda solves H lambda = ubar (H = dg/du, the Newton
// Hessian at u*, ONE 6x6 SPD solve per step) and databar = -lambda^T dg/d(data) in closed form.  Corner
// selection and the cone cases are piecewise constant (no gradient through them), as in autograd.
// Cost: about the arithmetic of three Newton visits per step, against 27 dual-number rollouts (cn_cube_tangent.cuh),
// which stay as the independent check of this derivation (tests/).
#pragma once
#include "cn_cube.cuh"

namespace cn {

// dPi/dy / eps of the cone projection at residual r (f = Pi(-r/eps)), symmetric [G00,G01,G02,G11,G12,G22]; the same
// three cases as cone_eval, without the friction scaling.
template <typename T> CN_HD void cone_jacobian(const T* r, T inv_eps, T* G) {
  const T t0 = -r[0] * inv_eps, t1 = -r[1] * inv_eps, n = -r[2] * inv_eps;
  const T rr2 = t0 * t0 + t1 * t1;
  const T rinv = t_rsqrt(t_max(rr2, t_tiny<T>()));
  const T rr = rr2 * rinv;
  const bool inside = rr <= n;
  const bool polar = (!inside) && (rr <= -n);
  const T s = T(0.5) * (n + rr);
  const T tx = t0 * rinv, ty = t1 * rinv;
  const T a = s * rinv, h = T(0.5);
  const T b00 = a * (T(1) - tx * tx) + h * tx * tx;
  const T b01 = (h - a) * tx * ty;
  const T b11 = a * (T(1) - ty * ty) + h * ty * ty;
  G[0] = (inside ? T(1) : (polar ? T(0) : b00)) * inv_eps;
  G[1] = ((inside || polar) ? T(0) : b01) * inv_eps;
  G[2] = ((inside || polar) ? T(0) : h * tx) * inv_eps;
  G[3] = (inside ? T(1) : (polar ? T(0) : b11)) * inv_eps;
  G[4] = ((inside || polar) ? T(0) : h * ty) * inv_eps;
  G[5] = (inside ? T(1) : (polar ? T(0) : h)) * inv_eps;
}

// Adjoint of quat_to_rot: Rb (3x3 row-major cotangent of R) -> += qb (4)
template <typename T> CN_HD void quat_to_rot_adjoint(const T* q, const T* Rb, T* qb) {
  const T w = q[0], x = q[1], y = q[2], z = q[3];
  const T n = w * w + x * x + y * y + z * z;
  const T s = T(2) / n;
  const T xs = x * s, ys = y * s, zs = z * s;
  const T yyb = -Rb[0] - Rb[8], zzb = -Rb[0] - Rb[4], xxb = -Rb[4] - Rb[8];
  const T xyb = Rb[1] + Rb[3], wzb = Rb[3] - Rb[1], xzb = Rb[2] + Rb[6], wyb = Rb[2] - Rb[6];
  const T yzb = Rb[5] + Rb[7], wxb = Rb[7] - Rb[5];
  T wb = wzb * zs + wyb * ys + wxb * xs;
  T xb = xzb * zs + xyb * ys + xxb * xs;
  T yb = yzb * zs + yyb * ys;
  T zb = zzb * zs;
  const T zsb = zzb * z + yzb * y + xzb * x + wzb * w;
  const T ysb = yyb * y + xyb * x + wyb * w;
  const T xsb = xxb * x + wxb * w;
  xb += xsb * s; yb += ysb * s; zb += zsb * s;
  const T sb = xsb * x + ysb * y + zsb * z;
  const T nb = -sb * s / n;
  qb[0] += wb + T(2) * w * nb;
  qb[1] += xb + T(2) * x * nb;
  qb[2] += yb + T(2) * y * nb;
  qb[3] += zb + T(2) * z * nb;
}

// Adjoint of the contact-free acceleration a = M^-1 F (cube_free_accel; state coordinates [w_body ; v_world]) and,
// with Fb = 0, of any quadratic form in the same mass matrix: given the cotangents Mb (6x6 row-major, of an
// unconstrained M, as autograd gives) and Fb (6) accumulates d/d inertia-vector into grad[0..9], and the state
// cotangents Rb (3x3), wb (3).
template <typename T>
CN_HD void cube_mass_force_adjoint(const CubeParams<T>& P, const T* R, const T* w, const T* Mb, const T* Fb, T* grad,
                                   T* Rb, T* wb) {
  T Kww[9], N[9], trvv = T(0);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Kww[3 * i + j] = Mb[6 * i + j];
      N[3 * i + j] = Mb[6 * i + 3 + j] + Mb[6 * (3 + j) + i];
    }
#pragma unroll
  for (int i = 0; i < 3; ++i) trvv += Mb[6 * (3 + i) + 3 + i];
  rigid_body_inertia_adjoint<T>(P.m, P.c, R, w, P.grav, Kww, N, trvv, Fb, grad);
  // R: M_wv = m S(c) R^T, M_vw = -m R S(c)  ->  Rb += m N^T S(c);  columns of S(c): S(c) e_j = c x e_j
  const T c0 = P.c[0], c1 = P.c[1], c2 = P.c[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const T n0 = N[i], n1 = N[3 + i], n2 = N[6 + i];        // row i of N^T
    // (N^T S(c))_{i,:} = [n0 n1 n2] S(c),  S(c) = [[0,-c2,c1],[c2,0,-c0],[-c1,c0,0]]
    Rb[3 * i + 0] += P.m * (n1 * c2 - n2 * c1);
    Rb[3 * i + 1] += P.m * (n2 * c0 - n0 * c2);
    Rb[3 * i + 2] += P.m * (n0 * c1 - n1 * c0);
  }
  // F_w = -w x (Io w) + m c x (R^T g), g = (0,0,-grav):  Rb[2][:] += -m grav (Fb_w x c)
  T fxc[3];
  cross3(Fb, P.c, fxc);
#pragma unroll
  for (int j = 0; j < 3; ++j) Rb[6 + j] += -P.m * P.grav * fxc[j];
  // F_v = -m R (w x (w x c)) + m g:  Rb += -m Fb_v (w x (w x c))^T
  T wc[3], wwc[3];
  cross3(w, P.c, wc); cross3(w, wc, wwc);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rb[3 * i + j] += -P.m * Fb[3 + i] * wwc[j];
  // w: from -w x (Io w) and -m R (w x (w x c))
  T Iw[3], t1[3], fxw[3], t2[3], b[3], t3[3], bxw[3], t4[3];
  sym3_mul(P.Io, w, Iw);
  cross3(Iw, Fb, t1);                    // (Io w) x Fb_w
  cross3(Fb, w, fxw);
  sym3_mul(P.Io, fxw, t2);               // Io (Fb_w x w)
  rot3t(R, Fb + 3, b);
  cross3(wc, b, t3);                     // (w x c) x b
  cross3(b, w, bxw);
  cross3(P.c, bxw, t4);                  // c x (b x w)
#pragma unroll
  for (int i = 0; i < 3; ++i) wb[i] += -t1[i] - t2[i] - P.m * (t3[i] + t4[i]);
}

// One step backwards.  x: the step's input state (13), u: its solved QP optimum (world twist, 6), xnb: cotangent of
// the step's output x+ (13).  Writes xb (13) = cotangent of x, and ADDS d/d[inertia 10 | mu | half 3] into grad.
template <typename T>
CN_HD void cube_step_backward(const CubeParams<T>& P, const T* x, const T* u, const T* xnb, T* xb, T* grad) {
  T store[CUBE_PROB_FIELDS];
  const CubeProb<T> S{store, 1};
  CubeStepAux<T> A;
  cube_step_prologue<T, 4>(P, x, S, A);                    // forward quantities: R, vm, sel, slot (IW, mcW, rho, q)
  const T* R = A.R;
  T acc[6];
  cube_free_accel(P, R, x + 7, acc, acc + 3);
  T uB[3], vn[6];
  rot3t(R, u, uB);
#pragma unroll
  for (int i = 0; i < 3; ++i) { vn[i] = A.vm[i] + uB[i]; vn[3 + i] = A.vm[3 + i] + u[3 + i]; }

  T Rb[9], qb[4], vnb[6], posb[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rb[i] = T(0);
  // ---- x+ = (quat (x) exp(vn_w dt), pos + vn_v dt, vn) ----
#pragma unroll
  for (int i = 0; i < 6; ++i) vnb[i] = xnb[7 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { posb[i] = xnb[4 + i]; vnb[3 + i] += P.dt * xnb[4 + i]; }
  {
    const T rx = vn[0] * P.dt, ry = vn[1] * P.dt, rz = vn[2] * P.dt;
    const T ang2 = rx * rx + ry * ry + rz * rz;
    const T ang = t_sqrt(ang2);
    const T h = T(0.5) * ang;
    const bool small = !(h > T(1e-4));
    using std::cos;
    using std::sin;
    const T sh = sin(h), ch = cos(h);
    const T sinc = small ? T(1) - h * h * (T(1) / T(6)) : sh / h;
    const T k = T(0.5) * sinc;
    const T dw = ch, dx = rx * k, dy = ry * k, dz = rz * k;
    const T qw = x[0], qx = x[1], qy = x[2], qz = x[3];
    const T b0 = xnb[0], b1 = xnb[1], b2 = xnb[2], b3 = xnb[3];
    qb[0] = b0 * dw + b1 * dx + b2 * dy + b3 * dz;
    qb[1] = -b0 * dx + b1 * dw - b2 * dz + b3 * dy;
    qb[2] = -b0 * dy + b1 * dz + b2 * dw - b3 * dx;
    qb[3] = -b0 * dz - b1 * dy + b2 * dx + b3 * dw;
    const T dwb = b0 * qw + b1 * qx + b2 * qy + b3 * qz;
    const T dxb = -b0 * qx + b1 * qw + b2 * qz - b3 * qy;
    const T dyb = -b0 * qy - b1 * qz + b2 * qw + b3 * qx;
    const T dzb = -b0 * qz + b1 * qy - b2 * qx + b3 * qw;
    const T kb = dxb * rx + dyb * ry + dzb * rz;
    // d h / d r = r / (4 h);  d dw / d h = -sin h;  d k / d h = (h cos h - sin h) / (2 h^2)
    const T coef = small ? (-dwb * T(0.25) * (T(1) - h * h * (T(1) / T(6))) - kb * (T(1) / T(24)))
                         : (-dwb * sh + kb * (h * ch - sh) / (T(2) * h * h)) / (T(4) * h);
    vnb[0] += P.dt * (k * dxb + coef * rx);
    vnb[1] += P.dt * (k * dyb + coef * ry);
    vnb[2] += P.dt * (k * dzb + coef * rz);
  }
  // ---- vn = [vm_w + R^T u_w ; vm_v + u_v] ----
  T vmb[6], ub[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) vmb[i] = vnb[i];
  rot3(R, vnb, ub);
#pragma unroll
  for (int i = 0; i < 3; ++i) ub[3 + i] = vnb[3 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rb[3 * i + j] += u[i] * vnb[j];
  // ---- implicit differentiation of the QP: H lambda = ubar at u* ----
  T g[6], H[36], res2, scale2, lam[6], nub[6];
  cube_eval<T, true, 4>(P, S, u, g, H, res2, scale2);
#pragma unroll
  for (int i = 0; i < 6; ++i) nub[i] = -ub[i];
  // Cholesky, not the adjugate block elimination of the Newton visits: lambda's tiny components along the stiff
  // (contact) directions are multiplied by 1/eps below, so the solve has to be backward stable
  chol_solve_neg<T, 6>(H, nub, lam);
  // world twists of vm
  T vmW[6], vmWb[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  rot3(R, A.vm, vmW);
#pragma unroll
  for (int i = 0; i < 3; ++i) vmW[3 + i] = A.vm[3 + i];
  T gmu = T(0), gh[3] = {T(0), T(0), T(0)};
  T posbz = T(0);
  const T inv_dt = T(1) / P.dt;
#pragma unroll
  for (int c = 0; c < CUBE_NC; ++c) {
    T rho[3], r[3], f[3], G[6];
    cube_contact_residual(P, S, c, u, rho, r);
    cone_eval<T, false>(r, P.inv_eps, P.mu, f, (T*)nullptr);
    cone_jacobian<T>(r, P.inv_eps, G);
    T e[3], wl[3];
    cross3(u, rho, e); cross3(lam, rho, wl);
#pragma unroll
    for (int i = 0; i < 3; ++i) { e[i] += u[3 + i]; wl[i] += lam[3 + i]; }
    // s = (G/eps) D_mu (J_c lambda)
    const T dw0 = P.mu * wl[0], dw1 = P.mu * wl[1], dw2 = wl[2];
    const T s0 = G[0] * dw0 + G[1] * dw1 + G[2] * dw2;
    const T s1 = G[1] * dw0 + G[3] * dw1 + G[4] * dw2;
    const T s2 = G[2] * dw0 + G[4] * dw1 + G[5] * dw2;
    const T ft[3] = {P.mu * f[0], P.mu * f[1], f[2]};
    const T ds[3] = {P.mu * s0, P.mu * s1, s2};
    T rhob[3], c1[3], c2[3];
    cross3(ft, lam, c1);                                   // ft x lambda_w
    cross3(ds, u, c2);                                     // (D_mu s) x u_w
#pragma unroll
    for (int i = 0; i < 3; ++i) rhob[i] = c1[i] - c2[i];
    gmu += (wl[0] * f[0] + wl[1] * f[1]) - (s0 * e[0] + s1 * e[1]);
    // q_c = D_mu (vmW_w x rho + vmW_v) + [0, 0, (rho_z + pos_z)/dt],  qbar_c = -s
    const T qb0 = -s0, qb1 = -s1, qb2 = -s2;
    const T a[3] = {P.mu * qb0, P.mu * qb1, qb2};
    T em[3], c3[3], c4[3];
    cross3(vmW, rho, em);
#pragma unroll
    for (int i = 0; i < 3; ++i) em[i] += vmW[3 + i];
    gmu += qb0 * em[0] + qb1 * em[1];
    cross3(rho, a, c3);                                    // vmWb_w += rho x a
    cross3(a, vmW, c4);                                    // rhob   += a x vmW_w
#pragma unroll
    for (int i = 0; i < 3; ++i) { vmWb[i] += c3[i]; vmWb[3 + i] += a[i]; rhob[i] += c4[i]; }
    rhob[2] += qb2 * inv_dt;
    posbz += qb2 * inv_dt;
    // rho_c = R (sigma_c o h)
    const T sg[3] = {sgn_bit<T>(A.sel, c, 0), sgn_bit<T>(A.sel, c, 1), sgn_bit<T>(A.sel, c, 2)};
    T rB[3];
    rot3t(R, rhob, rB);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gh[k] += sg[k] * rB[k];
#pragma unroll
      for (int i = 0; i < 3; ++i) Rb[3 * i + k] += rhob[i] * sg[k] * P.h[k];
    }
  }
  grad[10] += gmu;
  grad[11] += gh[0]; grad[12] += gh[1]; grad[13] += gh[2];
  // ---- mass terms of the QP: -lambda^T d(M^) u with M^ = T M T^T, T = blkdiag(R, I) ----
  T lamS[6], uS[6], Mb[36], zero6[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  rot3t(R, lam, lamS); rot3t(R, u, uS);
#pragma unroll
  for (int i = 0; i < 3; ++i) { lamS[3 + i] = lam[3 + i]; uS[3 + i] = u[3 + i]; }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) Mb[6 * i + j] = -lamS[i] * uS[j];
  T wb[3] = {T(0), T(0), T(0)};
  const T wzero[3] = {T(0), T(0), T(0)};
  cube_mass_force_adjoint<T>(P, R, wzero, Mb, zero6, grad, Rb, wb);      // Fb = 0: no force terms, w irrelevant
  {
    // the R-dependence through T: Rb += -lambda_w (M u_S)_w^T - u_w (M lambda_S)_w^T   (state-coordinate M)
    T MuW[6], MlW[6], MuS[3], MlS[3];
    cube_mass_mul(P, S, u, MuW); cube_mass_mul(P, S, lam, MlW);
    rot3t(R, MuW, MuS); rot3t(R, MlW, MlS);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rb[3 * i + j] += -lam[i] * MuS[j] - u[i] * MlS[j];
  }
  // ---- vmW = [R vm_w ; vm_v] ----
  T t3[3];
  rot3t(R, vmWb, t3);
#pragma unroll
  for (int i = 0; i < 3; ++i) { vmb[i] += t3[i]; vmb[3 + i] += vmWb[3 + i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rb[3 * i + j] += vmWb[i] * A.vm[j];
  // ---- vm = v + dt a(R, w) ----
  T ab[6], Fb[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) ab[i] = P.dt * vmb[i];
  cube_minv(P, R, ab, ab + 3, Fb, Fb + 3);                 // Fbar = M^-1 abar (M symmetric)
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) Mb[6 * i + j] = -Fb[i] * acc[j];
  cube_mass_force_adjoint<T>(P, R, x + 7, Mb, Fb, grad, Rb, wb);
  // ---- outputs ----
  quat_to_rot_adjoint<T>(x, Rb, qb);
#pragma unroll
  for (int i = 0; i < 4; ++i) xb[i] = qb[i];
  xb[4] = posb[0]; xb[5] = posb[1]; xb[6] = posb[2] + posbz;
#pragma unroll
  for (int i = 0; i < 3; ++i) { xb[7 + i] = vmb[i] + wb[i]; xb[10 + i] = vmb[3 + i]; }
}

// Backward of a whole rollout for one toss: traj (steps+1 states), usol (steps optima), xbar (cotangents of
// traj[1..steps]) -> gparams (14) and gx0 (13), both overwritten.
template <typename T>
CN_HD void cube_rollout_backward_sample(const CubeParams<T>& P, const T* traj, const T* usol, const T* xbar, int steps,
                                        T* gparams, T* gx0) {
  T cot[13], nxt[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) cot[i] = T(0);
#pragma unroll
  for (int i = 0; i < CUBE_NPARAM; ++i) gparams[i] = T(0);
  for (int s = steps - 1; s >= 0; --s) {
    T xs[13], us[6];
#pragma unroll
    for (int i = 0; i < 13; ++i) { xs[i] = traj[s * 13 + i]; cot[i] += xbar[s * 13 + i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) us[i] = usol[s * 6 + i];
    cube_step_backward<T>(P, xs, us, cot, nxt, gparams);
#pragma unroll
    for (int i = 0; i < 13; ++i) cot[i] = nxt[i];
  }
#pragma unroll
  for (int i = 0; i < 13; ++i) gx0[i] = cot[i];
}

}  // namespace cn
