// Learnable-parameter preparation on the device, and its reverse (chain rule), so that a
// loss + backward step is three launches with no PyTorch glue in between:
//   theta (log-Cholesky inertial parameters, inertia.py:46-66)  -> [m, c, I_cm / m]
//       InertialParameterConverter.theta_to_pi_o / pi_o_to_pi_cm / pi_cm_to_drake_spatial_inertia
//       (inertia.py:205-234, 304-331, 376-382), as LagrangianTerms.forward applies them
//       (multibody_terms.py:230-231)
//   friction_params -> |.| -> 2 mu_a mu_b / (mu_a + mu_b)        (multibody_terms.py:321-324, 466-471)
//   length_params   -> |.|                                       (geometry.py:394-397)
// The reverse direction is forward-mode: one thread per theta component propagates a dual
// number through the same templated map (10 x 10 Jacobian, negligible work).
#pragma once
#include "cn_common.cuh"

namespace cn {

template <typename T> struct Dual {
  T v, d;
  CN_HD Dual() : v(T(0)), d(T(0)) {}
  CN_HD Dual(T v_) : v(v_), d(T(0)) {}
  CN_HD Dual(T v_, T d_) : v(v_), d(d_) {}
};
template <typename T> CN_HD Dual<T> operator+(Dual<T> a, Dual<T> b) { return {a.v + b.v, a.d + b.d}; }
template <typename T> CN_HD Dual<T> operator-(Dual<T> a, Dual<T> b) { return {a.v - b.v, a.d - b.d}; }
template <typename T> CN_HD Dual<T> operator-(Dual<T> a) { return {-a.v, -a.d}; }
template <typename T> CN_HD Dual<T> operator*(Dual<T> a, Dual<T> b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <typename T> CN_HD Dual<T> operator/(Dual<T> a, Dual<T> b) {
  const T q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}
template <typename T> CN_HD Dual<T> t_exp(Dual<T> a) { const T e = exp(a.v); return {e, e * a.d}; }
CN_HD double t_exp(double a) { return exp(a); }
CN_HD float t_exp(float a) { return expf(a); }

// theta[10] = [alpha, d1, d2, d3, s12, s23, s13, t1, t2, t3]  ->  out[10] = [m, c(3), Ixx,Iyy,Izz,Ixy,Ixz,Iyz] / (m for I)
template <typename S> CN_HD void theta_to_inertia_vector(const S* th, S* out) {
  const S two(2), one(1);
  const S e1 = t_exp(th[1]), e2 = t_exp(th[2]), e3 = t_exp(th[3]);
  const S s12 = th[4], s23 = th[5], s13 = th[6], t1 = th[7], t2 = th[8], t3 = th[9];
  const S sc = t_exp(two * th[0]);
  const S m = sc * (t1 * t1 + t2 * t2 + t3 * t3 + one);
  const S h0 = sc * (t1 * e1), h1 = sc * (t1 * s12 + t2 * e2), h2 = sc * (t1 * s13 + t2 * s23 + t3 * e3);
  // inertia about the body origin (pi_o)
  const S oxx = sc * (s12 * s12 + s23 * s23 + s13 * s13 + e2 * e2 + e3 * e3);
  const S oyy = sc * (s13 * s13 + s23 * s23 + e1 * e1 + e3 * e3);
  const S ozz = sc * (s12 * s12 + e1 * e1 + e2 * e2);
  const S oxy = -(sc * (s12 * e1)), oxz = -(sc * (s13 * e1)), oyz = -(sc * (s12 * s13 + s23 * e2));
  const S c0 = h0 / m, c1 = h1 / m, c2 = h2 / m;
  // parallel axis, origin -> centre of mass:  I_cm = I_o - m (|c|^2 I - c c^T);  then / m
  out[0] = m; out[1] = c0; out[2] = c1; out[3] = c2;
  out[4] = oxx / m - (c1 * c1 + c2 * c2);
  out[5] = oyy / m - (c0 * c0 + c2 * c2);
  out[6] = ozz / m - (c0 * c0 + c1 * c1);
  out[7] = oxy / m + c0 * c1;
  out[8] = oxz / m + c0 * c2;
  out[9] = oyz / m + c1 * c2;
}

}  // namespace cn
