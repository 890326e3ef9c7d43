// newton.cuh -- device side of the MAP optimiser: Newton's method with
// the exact Hessian and step halving, the step before the samplers that provides `init`.
//
// Reference (paths relative to the reference root): Python/fit-jax.py:62-79 (same loop in
// fit-jax2.py:62-79 and fit-jax-ul.py:62-79):
//     for i in range(500):
//         g = glp(beta); step = -solve(hess(beta), g)
//         for j in range(15):
//             if lpost(beta+step) > lpost(beta): break
//             else: step = step/2
//         beta += step
//         if norm(g) < 0.01: break
// hess(lpost) = -(X' W X + diag(1/pscale^2)), W = diag(p_i (1 - p_i)), p = expit(X beta), so the
// step solves (X'WX + diag(pscale^-2)) step = glp(beta).  lpost and glp come from the fused
// evaluation kernel; this file adds the weights, the p x p contraction X'WX (64 x 64 blocks, SIMT
// float64 FMAs, per-CTA partial blocks summed in a fixed order), the Cholesky solve and the
// accept / halve decision, all on the device -- the host only sequences launches and reads one
// flag per trial.
#pragma once
#include "common.cuh"

namespace lrb {

constexpr int kHessPanel = 64;   // columns per Hessian block
constexpr int kHessRows = 32;    // rows per shared-memory tile

// w_i = p_i (1 - p_i), p_i = expit(x_i . beta); one thread per row, float64 arithmetic.
template <typename T>
__global__ void __launch_bounds__(kBlock) newton_weights_kernel(const T* __restrict__ X, long long n, int P, int p,
                                                                const double* __restrict__ beta, double* __restrict__ w) {
  __shared__ double sb[kMaxP];
  for (int j = threadIdx.x; j < P; j += kBlock) sb[j] = j < p ? beta[j] : 0.0;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * kBlock + threadIdx.x; row < n; row += (long long)gridDim.x * kBlock) {
    const T* x = X + row * P;
    double eta = 0.0;
    for (int j = 0; j < P; ++j) eta = fma((double)x[j], sb[j], eta);
    const double e = exp(-fabs(eta));            // overflow-free: sigma(|eta|) = 1/(1+e), sigma(-|eta|) = e/(1+e)
    const double inv = 1.0 / (1.0 + e);
    w[row] = (e * inv) * inv;                    // p (1 - p), symmetric in the sign of eta
  }
}

// One 64 x 64 block (rows ci0.., columns cj0.. of X'WX; PW = min(P, 64) of them are real) over a
// grid-strided set of 32-row tiles.  256 threads, thread (ti, tj) owns the 4 x 4 entries at
// (4 ti, 4 tj).  partial[blockIdx.x][64*64] receives this CTA's sum.
template <typename T>
__global__ void __launch_bounds__(kBlock) newton_hess_block_kernel(const T* __restrict__ X, const double* __restrict__ w,
                                                                   long long n, int P, int ci0, int cj0, int PW,
                                                                   double* __restrict__ partial) {
  __shared__ double sa[kHessRows][kHessPanel];   // w_r * x_r[ci0 + .]
  __shared__ double sb[kHessRows][kHessPanel];   // x_r[cj0 + .]
  const int tid = threadIdx.x;
  const int i0 = (tid / 16) * 4, j0 = (tid % 16) * 4;
  const bool active = i0 < PW && j0 < PW;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const long long ntiles = (n + kHessRows - 1) / kHessRows;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * kHessRows;
    for (int idx = tid; idx < kHessRows * PW; idx += kBlock) {
      const int r = idx / PW, c = idx % PW;
      const long long row = r0 + r;
      double xa = 0.0, xb = 0.0;
      if (row < n) {
        xa = w[row] * (double)X[row * P + ci0 + c];
        xb = (double)X[row * P + cj0 + c];
      }
      sa[r][c] = xa;
      sb[r][c] = xb;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < kHessRows; ++r) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = sa[r][i0 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = sb[r][j0 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
      }
    }
    __syncthreads();
  }
  if (active) {
    double* out = partial + (size_t)blockIdx.x * kHessPanel * kHessPanel;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) out[(i0 + a) * kHessPanel + j0 + b] = acc[a][b];
  }
}

// H[(ci0+i)*ld + cj0+j] (and its mirror image) = sum over CTAs, in CTA order, of partial[.][i*64+j].
// Diagonal blocks take their lower triangle only (fl(w x_i) x_j and fl(w x_j) x_i round differently),
// so H is exactly symmetric.
__global__ void newton_hess_reduce_kernel(const double* __restrict__ partial, int nparts, int PW, int ci0, int cj0, int p,
                                          double* __restrict__ H, int ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= PW * PW) return;
  const int i = idx / PW, j = idx % PW;
  if (ci0 + i >= p || cj0 + j >= p) return;
  if (ci0 == cj0 && j > i) return;
  double s = 0.0;
  for (int q = 0; q < nparts; ++q) s += partial[(size_t)q * kHessPanel * kHessPanel + i * kHessPanel + j];
  H[(size_t)(ci0 + i) * ld + cj0 + j] = s;
  H[(size_t)(cj0 + j) * ld + ci0 + i] = s;
}

// A = H + diag(1/pscale^2) (in place), A = L L' (Cholesky, lower, in place), solve A step = g.
// Plain serial float64 code shared by host and device so the arithmetic can be unit-tested
// without a GPU; p^3/3 flops (87 k at p = 64) on one thread is noise next to a pass over X.
// Returns 0, or k+1 if the matrix is not positive definite at column k.
__host__ __device__ inline int newton_chol_solve(double* A, int ld, int p, const double* pscale, const double* g,
                                                 double* step) {
  for (int j = 0; j < p; ++j) A[(size_t)j * ld + j] += 1.0 / (pscale[j] * pscale[j]);
  for (int k = 0; k < p; ++k) {
    double d = A[(size_t)k * ld + k];
    for (int m = 0; m < k; ++m) d -= A[(size_t)k * ld + m] * A[(size_t)k * ld + m];
    if (!(d > 0.0)) return k + 1;
    d = sqrt(d);
    A[(size_t)k * ld + k] = d;
    for (int i = k + 1; i < p; ++i) {
      double s = A[(size_t)i * ld + k];
      for (int m = 0; m < k; ++m) s -= A[(size_t)i * ld + m] * A[(size_t)k * ld + m];
      A[(size_t)i * ld + k] = s / d;
    }
  }
  for (int i = 0; i < p; ++i) {          // L y = g
    double s = g[i];
    for (int m = 0; m < i; ++m) s -= A[(size_t)i * ld + m] * step[m];
    step[i] = s / A[(size_t)i * ld + i];
  }
  for (int i = p - 1; i >= 0; --i) {     // L' s = y
    double s = step[i];
    for (int m = i + 1; m < p; ++m) s -= A[(size_t)m * ld + i] * step[m];
    step[i] = s / A[(size_t)i * ld + i];
  }
  return 0;
}

// Device-resident optimiser state.
struct NewtonState {
  double beta[kMaxP];      // current iterate
  double step[kMaxP];      // current (possibly halved) step
  double beta_try[kMaxP];  // beta + step: where the trial evaluation happens
  double lp_cur;           // lpost(beta)
  double grad_norm;        // ||glp(beta)||_2 at the iterate the step was computed from
  int32_t chol_fail;       // column+1 at which the Cholesky factorisation failed, else 0
  int32_t accepted;        // the last decision: 1 = trial point taken
  int32_t halvings;        // halvings done for the current step
  int32_t pad;
};

// res = [lpost, ll, lprior, glp...] of the evaluation at st->beta.  One thread.
__global__ void newton_step_kernel(NewtonState* st, double* H, int ld, int p, const double* pscale, const double* res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double gn = 0.0;
  for (int j = 0; j < p; ++j) gn += res[3 + j] * res[3 + j];
  st->grad_norm = sqrt(gn);
  st->lp_cur = res[0];
  st->chol_fail = newton_chol_solve(H, ld, p, pscale, res + 3, st->step);
  st->halvings = 0;
  st->accepted = 0;
  for (int j = 0; j < p; ++j) st->beta_try[j] = st->beta[j] + st->step[j];
}

// After the evaluation at beta_try (res[0] = its lpost): take it if it improves (fit-jax.py:71-72),
// else halve the step (:73-74).  force != 0: the 15 halvings are used up, take the step anyway (:75).
__global__ void newton_decide_kernel(NewtonState* st, int p, const double* res, int force) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool better = res[0] > st->lp_cur;
  if (better || force) {
    for (int j = 0; j < p; ++j) st->beta[j] = st->beta_try[j];
    st->accepted = 1;
  } else {
    for (int j = 0; j < p; ++j) {
      st->step[j] = st->step[j] / 2.0;
      st->beta_try[j] = st->beta[j] + st->step[j];
    }
    st->halvings += 1;
    st->accepted = 0;
  }
}

}  // namespace lrb
