// Forward-mode dual numbers with N tangent directions, so that the per-sample step code (templated on
// its scalar type) differentiates itself: x_next as a function of (state, parameters) carries
// d x_next / d (direction_k) through the prologue, the Newton solve, and the Lie-group update.
//
// Why forward mode is exact here: at a converged iterate u* the Newton update u - H^-1 g(u, theta)
// has tangent -H^-1 (dg/dtheta) dtheta (g ~ 0, dg/du = H), which is the implicit-function-theorem
// derivative of the QP solution -- what sappy's autograd backward returns to the reference
// (multibody_learnable_system.py:293-298, solver output NOT detached).  The solve is iterated in dual
// arithmetic, so the last steps (quadratically convergent, at least two inside the tolerance) carry
// exactly that tangent.  Comparisons act on the value only (piecewise-smooth code: cone cases, corner
// selection, |.|), as autograd does.
#pragma once
#include "cn_common.cuh"

namespace cn {

template <typename B, int N> struct DualN {
  B v;
  B d[N];
  CN_HD DualN() {}
  CN_HD DualN(B v_) : v(v_) {
    for (int i = 0; i < N; ++i) d[i] = B(0);
  }
  CN_HD DualN(int v_) : v(B(v_)) {
    for (int i = 0; i < N; ++i) d[i] = B(0);
  }
};

#define CN_DUAL_TPL template <typename B, int N>
#define CN_D DualN<B, N>

CN_DUAL_TPL CN_HD CN_D operator+(const CN_D& a, const CN_D& b) { CN_D r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D operator-(const CN_D& a, const CN_D& b) { CN_D r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D operator-(const CN_D& a) { CN_D r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D operator*(const CN_D& a, const CN_D& b) { CN_D r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D operator/(const CN_D& a, const CN_D& b) {
  CN_D r; const B ib = B(1) / b.v; r.v = a.v * ib;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
CN_DUAL_TPL CN_HD CN_D& operator+=(CN_D& a, const CN_D& b) { a = a + b; return a; }
CN_DUAL_TPL CN_HD CN_D& operator-=(CN_D& a, const CN_D& b) { a = a - b; return a; }
CN_DUAL_TPL CN_HD CN_D& operator*=(CN_D& a, const CN_D& b) { a = a * b; return a; }
CN_DUAL_TPL CN_HD bool operator<(const CN_D& a, const CN_D& b) { return a.v < b.v; }
CN_DUAL_TPL CN_HD bool operator<=(const CN_D& a, const CN_D& b) { return a.v <= b.v; }
CN_DUAL_TPL CN_HD bool operator>(const CN_D& a, const CN_D& b) { return a.v > b.v; }
CN_DUAL_TPL CN_HD bool operator>=(const CN_D& a, const CN_D& b) { return a.v >= b.v; }
CN_DUAL_TPL CN_HD bool operator!=(const CN_D& a, const CN_D& b) { return a.v != b.v; }
CN_DUAL_TPL CN_HD bool operator==(const CN_D& a, const CN_D& b) { return a.v == b.v; }

CN_DUAL_TPL CN_HD CN_D sqrt(const CN_D& a) {
  CN_D r; r.v = ::sqrt(a.v); const B k = r.v > B(0) ? B(0.5) / r.v : B(0);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k;
  return r;
}
CN_DUAL_TPL CN_HD double to_double(const CN_D& a) { return to_double(a.v); }
CN_DUAL_TPL CN_HD CN_D fabs(const CN_D& a) { return a.v < B(0) ? -a : a; }
CN_DUAL_TPL CN_HD CN_D sin(const CN_D& a) { CN_D r; r.v = ::sin(a.v); const B c = ::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D cos(const CN_D& a) { CN_D r; r.v = ::cos(a.v); const B s = -::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
CN_DUAL_TPL CN_HD CN_D t_rsqrt(const CN_D& a) {
  CN_D r; r.v = B(1) / ::sqrt(a.v); const B k = B(-0.5) * r.v / a.v;
  for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i];
  return r;
}
template <typename B, int N> struct Eps<DualN<B, N>> { static CN_HD DualN<B, N> v() { return DualN<B, N>(Eps<B>::v()); } };

#undef CN_DUAL_TPL
#undef CN_D

}  // namespace cn
