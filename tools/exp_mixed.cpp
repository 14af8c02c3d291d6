// Mixed-precision Newton solve of the cube loss QP on the CPU (the device math of csrc/cn_cube.cuh compiled for the host):
// visits per solve of the all-double solver against single-precision stage + double-precision finish (cube_solve_mixed), the
// histograms of both stages and the largest difference between the two optima.  Batch file: tools/exp_solver_trace.py
// writes it.  Build: g++ -O2 -std=c++17 -ffp-contract=off.  Development aid, not part of the package or the tests.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "../dair_pll_b200/csrc/cn_cube.cuh"
using namespace cn;
int main() {
  FILE* f = fopen("/tmp/dpll_exp_batch.bin", "rb");
  if (!f) { fprintf(stderr, "run tools/exp_solver_trace.py (it writes the batch file)\n"); return 1; }
  long B; size_t got = fread(&B, 8, 1, f);
  std::vector<double> x(13 * B), xp(13 * B); double inertia[10], mu[1], half[3];
  got += fread(x.data(), 8, 13 * B, f); got += fread(xp.data(), 8, 13 * B, f); got += fread(inertia, 8, 10, f); got += fread(mu, 8, 1, f);
  got += fread(half, 8, 3, f); fclose(f);
  if (got != (size_t)(1 + 26 * B + 14)) { fprintf(stderr, "short batch file\n"); return 1; }
  CubeParams<double> P; cube_params_init(P, inertia, mu, half, 0.0068, 1e-3);
  const SolverCfg<double> cfg = default_cfg<double>();
  long h32[128] = {0}, h64[128] = {0}, h0[128] = {0}; long t32 = 0, t64 = 0, t0 = 0, ns = 0; double maxdiff = 0;
  for (long b = 0; b < B; ++b) {
    double store[CUBE_PROB_FIELDS]; const CubeProb<double> S{store, 1}; CubeLossAux<double> A;
    cube_loss_prologue<double, 4>(P, &x[13 * b], &xp[13 * b], S, A);
    if (cube_trivially_solved<double, 4>(S)) continue;
    ++ns;
    double us[6], u0[6], u[6]; cube_loss_start<double>(A, (double)CN_LOSS_START_FACTOR, us);
    { double d[6], d0 = 0, best = -1; CubeTrial<double> tr{1, 0, 1}; int it = 0; long nv = 1;
      for (int i = 0; i < 6; ++i) u0[i] = us[i];
      while (cube_newton_visit<double, 4>(P, S, cfg, u0, d, d0, best, tr, it) != NEWTON_DONE) ++nv;
      h0[nv > 127 ? 127 : nv]++; t0 += nv; }
    for (int i = 0; i < 6; ++i) u[i] = us[i];
    int v[2]; cube_solve_mixed<double, 4>(P, S, cfg, u, v);
    h32[v[0] > 127 ? 127 : v[0]]++; h64[v[1] > 127 ? 127 : v[1]]++; t32 += v[0]; t64 += v[1];
    double n = 0, dd = 0; for (int i = 0; i < 6; ++i) { n += u0[i] * u0[i]; dd += (u[i] - u0[i]) * (u[i] - u0[i]); }
    if (n > 0 && sqrt(dd / n) > maxdiff) maxdiff = sqrt(dd / n);
  }
  printf("all double: %.2f visits/solve over %ld solves\n", (double)t0 / ns, ns);
  for (int i = 0; i < 128; ++i) if (h0[i]) printf("%d:%ld ", i, h0[i]);
  printf("\nmixed: single %.2f + double %.2f visits/solve, max rel diff of the optimum %.2e\nsingle: ", (double)t32 / ns, (double)t64 / ns, maxdiff);
  for (int i = 0; i < 128; ++i) if (h32[i]) printf("%d:%ld ", i, h32[i]);
  printf("\ndouble: ");
  for (int i = 0; i < 128; ++i) if (h64[i]) printf("%d:%ld ", i, h64[i]);
  printf("\n");
}
