/*
 * ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the friction-cone QP that the reference hands to the
 * un-vendored third-party solver `sappy.SAPSolver.apply(J, q, eps)`
 * (call sites: dair_pll/multibody_learnable_system.py:181-184 and :293-298;
 * problem statement in the docstring at :206-238).  sappy is pulled, unpinned,
 * from git+https://github.com/mshalm/sappy.git (setup.py:42) and is absent from
 * /root/reference, so this file restates the published problem
 *
 *     f* = argmin_{f in L3 x ... x L3}  1/2 f^T (A A^T + eps I) f + q^T f ,
 *     L3 = { (t_x, t_y, n) : n >= ||t||_2 }        ("sappy ordering", per contact)
 *
 * with A = J_M (k x n_v), k = 3 n_c (tensor_utils.py:460-497 fixes the ordering,
 * tensor_utils.py:393-458 fixes the cone).  PARITY UNPINNED at this boundary: the
 * reference holds no tests or golden vectors for the solver.  What makes the
 * answer checkable is that the problem is strictly convex (eps > 0), so the
 * optimum is unique and certified by the KKT conditions (f in K, s = Qf+q in K,
 * f^T s = 0), which tests/ verify independently of this code, and a second,
 * algorithmically unrelated solver (accelerated projected gradient on the dual,
 * oracle/cone_qp_apg below) agrees with it.
 *
 * Method (primal / SAP form): w in R^{n_v} minimises
 *     p(w) = 1/2 ||w||^2 + eps/2 sum_c || Pi( -(A_c w + q_c)/eps ) ||^2 ,
 * f_c = Pi(y_c); grad p = w - A^T f; hess p = I + A^T G A / eps.  Semismooth
 * Newton with an exact derivative-based line search (no function-value tests).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXC 24
#define MAXV 12
#define MAXK (3 * MAXC)

/* Projection of y=(t0,t1,n) onto L3 and (optionally) its Jacobian G (sym 3x3,
 * stored g00,g01,g02,g11,g12,g22). Same three cases as
 * tensor_utils.project_lorentz (tensor_utils.py:407-421). */
static void project_l3(const double y[3], double p[3], double G[6]) {
  double r = sqrt(y[0] * y[0] + y[1] * y[1]);
  double n = y[2];
  if (r <= n) {
    p[0] = y[0]; p[1] = y[1]; p[2] = y[2];
    if (G) { G[0] = 1; G[1] = 0; G[2] = 0; G[3] = 1; G[4] = 0; G[5] = 1; }
  } else if (r <= -n) {
    p[0] = p[1] = p[2] = 0.0;
    if (G) { G[0] = G[1] = G[2] = G[3] = G[4] = G[5] = 0.0; }
  } else {
    double s = 0.5 * (n + r);
    double tx = y[0] / r, ty = y[1] / r;
    p[0] = s * tx; p[1] = s * ty; p[2] = s;
    if (G) {
      double a = s / r;
      G[0] = a * (1.0 - tx * tx) + 0.5 * tx * tx;
      G[1] = (0.5 - a) * tx * ty;
      G[2] = 0.5 * tx;
      G[3] = a * (1.0 - ty * ty) + 0.5 * ty * ty;
      G[4] = 0.5 * ty;
      G[5] = 0.5;
    }
  }
}

/* dphi/dalpha and d2phi/dalpha2 of phi(alpha) = p(w + alpha d). */
static void line_derivs(int nc, int nv, const double* w, const double* d,
                        const double* r0, const double* Ad, double eps,
                        double alpha, double* d1, double* d2) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < nv; ++i) { a += (w[i] + alpha * d[i]) * d[i]; b += d[i] * d[i]; }
  for (int c = 0; c < nc; ++c) {
    double y[3], p[3], G[6];
    for (int j = 0; j < 3; ++j) y[j] = -(r0[3 * c + j] + alpha * Ad[3 * c + j]) / eps;
    project_l3(y, p, G);
    const double* e = Ad + 3 * c;
    a -= e[0] * p[0] + e[1] * p[1] + e[2] * p[2];
    b += (e[0] * (G[0] * e[0] + G[1] * e[1] + G[2] * e[2]) +
          e[1] * (G[1] * e[0] + G[3] * e[1] + G[4] * e[2]) +
          e[2] * (G[2] * e[0] + G[4] * e[1] + G[5] * e[2])) / eps;
  }
  *d1 = a; *d2 = b;
}

/* One sample. A: k x nv row-major. Returns iterations used. */
static int solve_one(int nc, int nv, const double* A, const double* q, double eps,
                     int max_iter, double tol, double ls_tol, const double* w0,
                     double* f, double* w_out, double* resid_out) {
  const int k = 3 * nc;
  double w[MAXV], g[MAXV], d[MAXV], r0[MAXK], Ad[MAXK], G[MAXC][6];
  double H[MAXV][MAXV], L[MAXV][MAXV];
  for (int i = 0; i < nv; ++i) w[i] = w0 ? w0[i] : 0.0;
  int it = 0;
  double best = INFINITY; int stall = 0;
  for (;; ++it) {
    /* residual r0 = A w + q, forces, gradient */
    double nw = 0, naf = 0, ng = 0;
    for (int i = 0; i < k; ++i) {
      double s = q[i];
      for (int j = 0; j < nv; ++j) s += A[i * nv + j] * w[j];
      r0[i] = s;
    }
    for (int c = 0; c < nc; ++c) {
      double y[3];
      for (int j = 0; j < 3; ++j) y[j] = -r0[3 * c + j] / eps;
      project_l3(y, f + 3 * c, G[c]);
    }
    for (int j = 0; j < nv; ++j) {
      double s = 0;
      for (int i = 0; i < k; ++i) s += A[i * nv + j] * f[i];
      g[j] = w[j] - s;
      nw += w[j] * w[j]; naf += s * s; ng += g[j] * g[j];
    }
    double scale = sqrt(nw > naf ? nw : naf);
    double res = sqrt(ng);
    *resid_out = scale > 0 ? res / scale : res;
    if (res <= tol * scale || res == 0.0) break;
    if (it >= max_iter) break;
    /* stall detection: residual has hit the rounding floor */
    if (res < best * 0.5) { best = res; stall = 0; }
    else if (++stall >= 3 && res <= 1e-10 * scale) break;
    /* H = I + A^T G A / eps */
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) H[i][j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < nc; ++c) {
      const double* Gc = G[c];
      const double* A0 = A + (3 * c + 0) * nv;
      const double* A1 = A + (3 * c + 1) * nv;
      const double* A2 = A + (3 * c + 2) * nv;
      for (int i = 0; i < nv; ++i) {
        double ga0 = (Gc[0] * A0[i] + Gc[1] * A1[i] + Gc[2] * A2[i]) / eps;
        double ga1 = (Gc[1] * A0[i] + Gc[3] * A1[i] + Gc[4] * A2[i]) / eps;
        double ga2 = (Gc[2] * A0[i] + Gc[4] * A1[i] + Gc[5] * A2[i]) / eps;
        for (int j = 0; j < nv; ++j) H[i][j] += ga0 * A0[j] + ga1 * A1[j] + ga2 * A2[j];
      }
    }
    /* Cholesky H = L L^T, d = -H^{-1} g */
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j <= i; ++j) {
        double s = H[i][j];
        for (int m = 0; m < j; ++m) s -= L[i][m] * L[j][m];
        L[i][j] = (i == j) ? sqrt(s) : s / L[j][j];
      }
    for (int i = 0; i < nv; ++i) {
      double s = -g[i];
      for (int m = 0; m < i; ++m) s -= L[i][m] * d[m];
      d[i] = s / L[i][i];
    }
    for (int i = nv - 1; i >= 0; --i) {
      double s = d[i];
      for (int m = i + 1; m < nv; ++m) s -= L[m][i] * d[m];
      d[i] = s / L[i][i];
    }
    for (int i = 0; i < k; ++i) {
      double s = 0;
      for (int j = 0; j < nv; ++j) s += A[i * nv + j] * d[j];
      Ad[i] = s;
    }
    /* exact line search on phi'(alpha) = 0 (monotone increasing) */
    double d0, h0, d1, h1;
    line_derivs(nc, nv, w, d, r0, Ad, eps, 0.0, &d0, &h0);
    double alpha = 1.0;
    line_derivs(nc, nv, w, d, r0, Ad, eps, 1.0, &d1, &h1);
    double ltol = ls_tol * fabs(d0);
    if (fabs(d1) > ltol) {
      double lo = 0.0, hi = 1.0, dhi = d1;
      int guard = 0;
      while (dhi < 0 && guard++ < 60) { lo = hi; hi *= 2.0; line_derivs(nc, nv, w, d, r0, Ad, eps, hi, &dhi, &h1); }
      /* rtsafe: Newton on phi' safeguarded by bisection in [lo,hi] */
      alpha = 0.5 * (lo + hi);
      for (int ls = 0; ls < 100; ++ls) {
        double da, ha;
        line_derivs(nc, nv, w, d, r0, Ad, eps, alpha, &da, &ha);
        if (fabs(da) <= ltol) break;
        if (da < 0) lo = alpha; else hi = alpha;
        double an = alpha - da / ha;
        if (!(an > lo && an < hi)) an = 0.5 * (lo + hi);
        if (hi - lo <= 1e-16 * hi || an == alpha) { alpha = an; break; }
        alpha = an;
      }
    }
    for (int i = 0; i < nv; ++i) w[i] += alpha * d[i];
  }
  for (int i = 0; i < nv; ++i) w_out[i] = w[i];
  return it;
}

/* Batched entry point. A (B,k,nv), q (B,k) in sappy ordering -> f (B,k), w (B,nv).
 * ls_tol: line search stops at |phi'(a)| <= ls_tol |phi'(0)| (1e-14 = exact);
 * w0: optional (B,nv) warm start (NULL = 0). */
int cone_qp_solve_batch(const double* A, const double* q, double eps, int64_t B,
                        int nc, int nv, int max_iter, double tol, double ls_tol,
                        const double* w0, int nthreads,
                        double* f, double* w, int32_t* iters, double* resid) {
  if (nc > MAXC || nv > MAXV || nc < 0 || nv <= 0) return 1;
  const int k = 3 * nc;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t b = 0; b < B; ++b) {
    double rr;
    int it = solve_one(nc, nv, A + b * k * nv, q + b * k, eps, max_iter, tol, ls_tol,
                       w0 ? w0 + b * nv : 0, f + b * k, w + b * nv, &rr);
    if (iters) iters[b] = it;
    if (resid) resid[b] = rr;
  }
  return 0;
}

/*
 * Independent cross-check solver: accelerated projected gradient (FISTA with
 * function-free adaptive restart) on the DUAL problem with explicit
 * Q = A A^T + eps I.  Shares only project_l3 with the Newton solver above.
 * Slow (O(sqrt(cond)) iterations) -- used on small golden sets only.
 */
int cone_qp_apg_batch(const double* A, const double* q, double eps, int64_t B,
                      int nc, int nv, int max_iter, double tol, double* f,
                      int32_t* iters) {
  if (nc > MAXC || nv > MAXV) return 1;
  const int k = 3 * nc;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t b = 0; b < B; ++b) {
    const double* Ab = A + b * k * nv;
    const double* qb = q + b * k;
    double Q[MAXK][MAXK], x[MAXK], z[MAXK], xn[MAXK], gr[MAXK];
    double tr = 0;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) {
        double s = (i == j) ? eps : 0.0;
        for (int m = 0; m < nv; ++m) s += Ab[i * nv + m] * Ab[j * nv + m];
        Q[i][j] = s;
        if (i == j) tr += s;
      }
    /* power iteration for lambda_max (upper bound by trace as fallback) */
    double v[MAXK], lam = tr;
    for (int i = 0; i < k; ++i) v[i] = 1.0 + 0.01 * i;
    for (int itp = 0; itp < 200; ++itp) {
      double u[MAXK], nn = 0;
      for (int i = 0; i < k; ++i) { double s = 0; for (int j = 0; j < k; ++j) s += Q[i][j] * v[j]; u[i] = s; nn += s * s; }
      nn = sqrt(nn);
      for (int i = 0; i < k; ++i) v[i] = u[i] / nn;
      lam = nn;
    }
    double Lc = 1.02 * lam + 1e-300, t = 1.0, qnorm = 0;
    for (int i = 0; i < k; ++i) qnorm += qb[i] * qb[i];
    qnorm = sqrt(qnorm);
    for (int i = 0; i < k; ++i) x[i] = z[i] = 0.0;
    int it = 0;
    for (; it < max_iter; ++it) {
      for (int i = 0; i < k; ++i) { double s = qb[i]; for (int j = 0; j < k; ++j) s += Q[i][j] * z[j]; gr[i] = s; }
      for (int c = 0; c < nc; ++c) {
        double y[3];
        for (int j = 0; j < 3; ++j) y[j] = z[3 * c + j] - gr[3 * c + j] / Lc;
        project_l3(y, xn + 3 * c, 0);
      }
      /* gradient-mapping restart test + convergence on fixed-point residual */
      double dot = 0, dn = 0, xnorm = 0;
      for (int i = 0; i < k; ++i) { dot += (z[i] - xn[i]) * (xn[i] - x[i]); dn += (xn[i] - z[i]) * (xn[i] - z[i]); xnorm += xn[i] * xn[i]; }
      double tn = 0.5 * (1.0 + sqrt(1.0 + 4.0 * t * t));
      double beta = (t - 1.0) / tn;
      if (dot > 0) { tn = 1.0; beta = 0.0; }
      for (int i = 0; i < k; ++i) { z[i] = xn[i] + beta * (xn[i] - x[i]); x[i] = xn[i]; }
      t = tn;
      if (sqrt(dn) * Lc <= tol * (sqrt(xnorm) * lam + qnorm) + 1e-300) { ++it; break; }
    }
    for (int i = 0; i < k; ++i) f[b * k + i] = x[i];
    if (iters) iters[b] = it;
  }
  return 0;
}
