// Newton-visit traces of the cube loss QP on the CPU (the device math of csrc/cn_cube.cuh compiled for the host): per visit the
// cone case of every contact (S inside / B boundary / . polar), its (n, |t|) in units of the force, the relative residual and
// the line-search state; then the histogram of visits per solve over the batch.  Driver: tools/exp_solver_trace.py.
// Development aid (how the long chains look), not part of the package or the tests.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
static long g_visits = 0;
#define CN_STAT_LS()
#define CN_STAT_UNIT() (++g_visits)
#include "../dair_pll_b200/csrc/cn_cube.cuh"
using namespace cn;
// reads raw doubles: x (B,13), xp (B,13), inertia 10, mu 1, half 3
int main(int argc, char** argv) {
  FILE* f = fopen(argc > 3 ? argv[3] : "/tmp/dpll_exp_batch.bin", "rb");
  if (!f) { fprintf(stderr, "run tools/exp_solver_trace.py (it writes the batch file)\n"); return 1; }
  long B; size_t got = fread(&B, 8, 1, f);
  std::vector<double> x(13 * B), xp(13 * B); double inertia[10], mu[1], half[3];
  got += fread(x.data(), 8, 13 * B, f); got += fread(xp.data(), 8, 13 * B, f); got += fread(inertia, 8, 10, f); got += fread(mu, 8, 1, f);
  got += fread(half, 8, 3, f);
  if (got != (size_t)(1 + 26 * B + 14)) { fprintf(stderr, "short batch file\n"); return 1; }
  fclose(f);
  CubeParams<double> P; cube_params_init(P, inertia, mu, half, 0.0068, 1e-3);
  SolverCfg<double> cfg = default_cfg<double>();
  int target = argc > 1 ? atoi(argv[1]) : 10, nshow = argc > 2 ? atoi(argv[2]) : 3;
  long hist[128] = {0}; long tot = 0, ns = 0;
  for (long b = 0; b < B; ++b) {
    double store[CUBE_PROB_FIELDS]; const CubeProb<double> S{store, 1}; CubeLossAux<double> A;
    cube_loss_prologue<double, 4>(P, &x[13 * b], &xp[13 * b], S, A);
    if (cube_trivially_solved<double, 4>(S)) { hist[0]++; continue; }
    const bool zero = getenv("ZERO_START") != nullptr;      // the round-1 start point u = 0, for comparison
    double u[6], d[6], d0 = 0, best = -1; CubeTrial<double> tr{1, 0, 1}; int it = 0;
    cube_loss_start<double>(A, zero ? 0.0 : (double)CN_LOSS_START_FACTOR, u);
    long v0 = g_visits;
    bool show = false;
    // dry run to get count
    {
      double u2[6], d2[6], d02 = 0, best2 = -1; CubeTrial<double> tr2{1, 0, 1}; int it2 = 0;
      for (int i = 0; i < 6; ++i) u2[i] = u[i];
      while (cube_newton_visit<double, 4>(P, S, cfg, u2, d2, d02, best2, tr2, it2) != NEWTON_DONE) {}
      long nv = g_visits - v0; hist[nv > 127 ? 127 : nv]++; tot += nv; ns++;
      show = (nv == target && nshow > 0);
    }
    if (!show) continue;
    nshow--;
    printf("sample %ld  q:", b);
    for (int c = 0; c < 4; ++c) printf(" [%.3e %.3e %.3e]", S.q(3*c), S.q(3*c+1), S.q(3*c+2));
    printf("\n");
    int rc;
    do {
      // modes at current u
      double g[6], H[36], res2, sc2;
      cube_eval<double, true, 4>(P, S, u, g, H, res2, sc2);
      char modes[5] = {0};
      for (int c = 0; c < 4; ++c) { double rho[3], r[3]; cube_contact_residual(P, S, c, u, rho, r);
        double t0 = -r[0], t1 = -r[1], n = -r[2]; double rr = sqrt(t0*t0+t1*t1); modes[c] = rr <= n ? 'S' : (rr <= -n ? '.' : 'B'); }
      rc = cube_newton_visit<double, 4>(P, S, cfg, u, d, d0, best, tr, it);
      for (int c = 0; c < 4; ++c) { double rho[3], r[3]; cube_contact_residual(P, S, c, u, rho, r);
        double t0 = -r[0]*P.inv_eps, t1 = -r[1]*P.inv_eps, n = -r[2]*P.inv_eps; double rr = sqrt(t0*t0+t1*t1); printf("   c%d n %.3e rr %.3e ang %.2f |", c, n, rr, atan2(t1,t0)); }
      printf("\n");
      printf("  modes %s relres %.2e  it %d trials %d alpha %.3f\n", modes, sqrt(res2 / sc2), it & 0xff, (it >> 8) & 0xff, tr.alpha);
    } while (rc != NEWTON_DONE);
  }
  printf("visits/solve %.2f  solves %ld\n", (double)tot / ns, ns);
  for (int i = 0; i < 128; ++i) if (hist[i]) printf("%d:%ld ", i, hist[i]);
  printf("\n");
}
