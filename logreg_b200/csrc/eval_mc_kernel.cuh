// eval_mc_kernel.cuh -- SIMT many-chain variant of the fused kernel: each 8 KB batch
// of X is loaded ONCE into registers and reused for NC chains (coefficient vectors),
// so X traffic per chain drops NC-fold.  This is the path for small chain counts;
// above the chain-count threshold the tcgen05 kernel (eval_tc_kernel.cuh) takes
// over, where eta = X B' and G = R' X become tensor-core contractions.
//
// Same arithmetic per chain as eval_kernel (see eval_kernel.cuh for the lane
// mapping); the finish step runs in a separate launch with one CTA per chain.
#pragma once
#include "eval_kernel.cuh"

namespace lrb {

constexpr int kSumStride = kMaxP + 1;   // doubles per chain in the sums array
constexpr int kResStride = kMaxP + 3;   // doubles per chain in the result array

struct EvalMcArgs {
  const void* X;
  const uint8_t* y;
  long long n;
  double* partials;          // [grid][NC][P+1]
  unsigned int* ticket;
  const double* beta_base;   // chain c's coefficients at beta_base + c*beta_stride
  long long beta_stride;     // in doubles
  int chain0, nc_active, p;
  double* sums;              // [C][kSumStride]: [ll, gll] per chain
  const SamplerState* states;  // nullptr for a bare evaluation (else: pause check)
};

template <typename T, int P, int NC>
__global__ void __launch_bounds__(kBlock, 1) eval_mc_kernel(const EvalMcArgs a) {
  using C = Chunk<T>;
  using vec = typename C::vec;
  using A = typename C::acc_t;
  constexpr int V = C::V;
  constexpr int CPR = P / V;
  constexpr int L = CPR < 32 ? CPR : 32;
  constexpr int SPR = CPR / L;
  constexpr int G = 32 / L;
  constexpr int S = 16;
  constexpr int SG = S / SPR;
  constexpr int RB = SG * G;
  constexpr int LOG_L = ilog2(L), LOG_SG = ilog2(SG);
  constexpr int NH = LOG_L < LOG_SG ? LOG_L : LOG_SG;
  constexpr int M = SG >> NH;

  __shared__ double red[kWarps][P];
  __shared__ double redll[kWarps];
  __shared__ unsigned int s_ticket;

  if (a.states != nullptr && a.states[a.chain0].phase == PH_PAUSED) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_in_row = lane & (L - 1);
  const int g_in_slab = lane >> LOG_L;
  const vec* __restrict__ Xv = reinterpret_cast<const vec*>(a.X);
  const long long n = a.n;
  const long long total_chunks = n * CPR;
  const long long nbatch = (n + RB - 1) / RB;

  A bh[NC][SPR][V];
  float bl[NC][SPR][V];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int cc = c < a.nc_active ? c : 0;   // surplus slots recompute chain0's first chain (discarded)
    const double* beta = a.beta_base + (long long)(a.chain0 + cc) * a.beta_stride;
#pragma unroll
    for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int col = (sl * L + t_in_row) * V + i;
        const double b = col < a.p ? beta[col] : 0.0;
        bh[c][sl][i] = (A)b;
        bl[c][sl][i] = (float)(b - (double)bh[c][sl][i]);
      }
  }

  int khigh = 0;
#pragma unroll
  for (int b = 0; b < NH; ++b) khigh += ((t_in_row >> b) & 1) * (SG >> (b + 1));
  const bool owner = (t_in_row >> NH) == 0;

  double acc[NC][SPR][V];
  double ll_acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    ll_acc[c] = 0.0;
#pragma unroll
    for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[c][sl][i] = 0.0;
  }

  const long long nwarps = (long long)gridDim.x * kWarps;
  for (long long bt = (long long)blockIdx.x * kWarps + warp; bt < nbatch; bt += nwarps) {
    const long long chunk0 = bt * (S * 32) + lane;
    const long long row0 = bt * RB;
    vec v[S];
    if (row0 + RB <= n) {
#pragma unroll
      for (int s = 0; s < S; ++s) v[s] = ldg_stream(Xv + chunk0 + s * 32);
    } else {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const long long c = chunk0 + s * 32;
        v[s] = c < total_chunks ? ldg_stream(Xv + c) : zero_vec((vec*)nullptr);
      }
    }
    bool y1[M], valid[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const long long row = row0 + (long long)(j + khigh) * G + g_in_slab;
      valid[j] = owner && row < n;
      y1[j] = valid[j] ? (a.y[row] != 0) : false;
    }

#pragma unroll
    for (int c = 0; c < NC; ++c) {
      A q[SG];
#pragma unroll
      for (int k = 0; k < SG; ++k) {
        A s_hi = (A)0;
#pragma unroll
        for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
          for (int i = 0; i < V; ++i) s_hi = fma(elem(v[k * SPR + sl], i), bh[c][sl][i], s_hi);
        if constexpr (sizeof(T) == 4) {
          float s_lo = 0.f;
#pragma unroll
          for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
            for (int i = 0; i < V; ++i) s_lo = fmaf((float)elem(v[k * SPR + sl], i), bl[c][sl][i], s_lo);
          s_hi += (A)s_lo;
        }
        q[k] = s_hi;
      }
      reduce_rows<SG, 0, LOG_L>(q, lane);
      A r[M];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        A rr;
        const A lt = row_terms(q[j], y1[j], rr);
        r[j] = rr;
        if (valid[j]) ll_acc[c] += (double)lt;
      }
      A gb[SPR][V];
#pragma unroll
      for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
        for (int i = 0; i < V; ++i) gb[sl][i] = (A)0;
#pragma unroll
      for (int k = 0; k < SG; ++k) {
        const int j = k & (M - 1);
        int tsrc = 0;
#pragma unroll
        for (int b = 0; b < NH; ++b) tsrc |= ((k >> (LOG_SG - 1 - b)) & 1) << b;
        const A rr = __shfl_sync(0xffffffffu, r[j], (lane & ~(L - 1)) | tsrc);
#pragma unroll
        for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
          for (int i = 0; i < V; ++i) gb[sl][i] = fma(rr, (A)elem(v[k * SPR + sl], i), gb[sl][i]);
      }
#pragma unroll
      for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[c][sl][i] += (double)gb[sl][i];
    }
  }

  // ---- CTA reduction, one chain at a time through the same scratch
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int sl = 0; sl < SPR; ++sl)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        double x = acc[c][sl][i];
#pragma unroll
        for (int o = L; o < 32; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane < L) red[warp][(sl * L + lane) * V + i] = x;
      }
    const double lw = warp_sum(ll_acc[c]);
    if (lane == 0) redll[warp] = lw;
    __syncthreads();
    double* mypart = a.partials + ((size_t)blockIdx.x * NC + c) * (P + 1);
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += redll[w];
      mypart[0] = s;
    }
    for (int col = tid; col < P; col += kBlock) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[w][col];
      mypart[1 + col] = s;
    }
    __syncthreads();
  }

  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  if (tid == 0) *a.ticket = 0u;

  // last CTA: fixed-order sum of the CTA partials for each active chain
  for (int idx = tid; idx < a.nc_active * (P + 1); idx += kBlock) {
    const int c = idx / (P + 1), col = idx % (P + 1);
    const double* src = a.partials + (size_t)c * (P + 1) + col;
    const size_t stride = (size_t)NC * (P + 1);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = 0;
    for (; b + 3 < (int)gridDim.x; b += 4) {
      const double v0 = __ldcg(src + (size_t)b * stride), v1 = __ldcg(src + (size_t)(b + 1) * stride);
      const double v2 = __ldcg(src + (size_t)(b + 2) * stride), v3 = __ldcg(src + (size_t)(b + 3) * stride);
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; b < (int)gridDim.x; ++b) s0 += __ldcg(src + (size_t)b * stride);
    if (col <= a.p) a.sums[(size_t)(a.chain0 + c) * kSumStride + col] = (s0 + s1) + (s2 + s3);
  }
}

// One CTA per chain: prior + result + sampler update from the reduced sums.
__global__ void finish_mc_kernel(FinishArgs base, const double* sums, SamplerState* states,
                                 const double* beta_base, long long beta_stride, double* res) {
  __shared__ double s_sums[kMaxP + 1];
  __shared__ double scratch[kWarps];
  const int c = blockIdx.x;
  FinishArgs f = base;
  f.state = states ? states + c : nullptr;
  if (f.state && f.state->phase == PH_PAUSED) return;
  f.beta = beta_base + (long long)c * beta_stride;
  f.res = res + (size_t)c * kResStride;
  f.p2p = 0;
  for (int i = threadIdx.x; i <= f.p; i += kBlock) s_sums[i] = sums[(size_t)c * kSumStride + i];
  __syncthreads();
  finish_eval(f, s_sums, scratch);
}

}  // namespace lrb
