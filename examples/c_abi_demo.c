/* Plain-C use of the library (no Python, no torch): evaluate lpost / glp and run a short HMC chain
 * on a tiny problem.  Build:  gcc -std=c99 -Iinclude examples/c_abi_demo.c -o c_abi_demo \
 *                                 -Llogreg_b200/_lib -llogreg_b200 -Wl,-rpath,$PWD/logreg_b200/_lib -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "logreg_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != LRB_OK) {                                                     \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lrb_last_error(h));      \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  enum { N = 1000, P = 4 };
  static double X[N * P]; /* row-major */
  static float y[N];
  double pscale[P] = {10.0, 1.0, 1.0, 1.0}, beta[P] = {0.1, -0.2, 0.3, 0.0};
  double lpost, ll, glp[P], init[P] = {0, 0, 0, 0}, scale[P] = {1, 1, 1, 1}, out[50 * P];
  int64_t accepted = 0;
  lrb_handle* h = NULL;
  lrb_sampler_params sp;
  unsigned s = 12345u;
  int i, j;
  for (i = 0; i < N; ++i) {
    X[i * P] = 1.0;
    for (j = 1; j < P; ++j) { s = s * 1664525u + 1013904223u; X[i * P + j] = (double)(s >> 8) / 8388608.0 - 1.0; }
    s = s * 1664525u + 1013904223u;
    y[i] = (float)((s >> 16) & 1u);
  }
  CHECK(lrb_create(0, &h));
  CHECK(lrb_bind_data(h, X, LRB_F64, LRB_ROW_MAJOR, P, y, LRB_F32, N, P, pscale, LRB_MODE_FP64, LRB_HOST));
  CHECK(lrb_eval(h, beta, 1, 1, &lpost, &ll, glp));
  printf("lpost %.10f ll %.10f glp[0] %.10f\n", lpost, ll, glp[0]);
  sp.sampler = LRB_HMC; sp.l = 10; sp.step = 0.02; sp.scale = scale; sp.seed = 7; sp.rng = LRB_RNG_PHILOX;
  sp.flags = 0; sp.init_lpost = -INFINITY; sp.t0 = 0;
  CHECK(lrb_run(h, &sp, init, 1, 2, 50, NULL, NULL, out, &accepted));
  printf("HMC: 50 thinned samples, accepted %lld of 100, last state %.4f %.4f %.4f %.4f\n", (long long)accepted,
         out[49 * P], out[49 * P + 1], out[49 * P + 2], out[49 * P + 3]);
  lrb_destroy(h);
  return 0;
}
