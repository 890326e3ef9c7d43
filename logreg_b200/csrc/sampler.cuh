// sampler.cuh -- on-device sampler state machine (RWMH / UL / MALA / HMC) and the
// "finish" step that turns the reduced row sums into lpost / glp.
//
// One CTA of kBlock threads executes these functions; thread j owns coefficient j.
// Every reduction over coefficients is a fixed-order block sum, so all ranks of a
// row-sharded run (same inputs after the allreduce, same code) take bit-identical
// decisions without any broadcast.
//
// Reference semantics followed (paths relative to the reference root):
//   RWMH  Python/fit-numpy.py:53-62 (mhKernel), :81-84 (rprop), :64-79 (mcmc)
//   UL    Python/fit-np-ul.py:61-68 (ulKernel), :70-84 (mcmc)
//   MALA  Python/fit-np-mala.py:61-70 (mhKernel), :72-78 (malaKernel), :80-95 (mcmc)
//   HMC   Python/fit-np-hmc.py:56-63 (mhKernel), :65-87 (hmcKernel), :89-103 (mcmc)
// What differs, by design: the reference re-evaluates glp/lpost at the current
// state on every step (7 passes over X per MALA step, 2l+4 per HMC step); here
// (lpost, glp) of the current state are cached, so MALA costs ONE fused pass and
// HMC costs l.  The MALA acceptance ratio uses the closed
// form in which the Gaussian normalisers cancel (Dex/fit-mala.dx:65-68).
#pragma once
#include "common.cuh"

namespace lrb {

enum { PH_PAUSED = 0, PH_INIT = 1, PH_STEP = 2 };
enum { S_RWMH = 0, S_UL = 1, S_MALA = 2, S_HMC = 3 };
enum { RNG_PHILOX = 0, RNG_REPLAY = 1 };

struct SamplerState {
  int32_t kind, l, p, rng;
  double step, sqrt_step;
  uint64_t seed;
  long long t;          // kernel applications completed since the chain began
  long long t_run0;     // t when the current run began (thinning rows count from here)
  long long t_replay0;  // t at which the replay arrays start
  long long t_end;      // pause when t reaches this
  long long thin;
  long long accepted;   // accepted proposals since the chain began
  int32_t phase, leap;
  double lp_x;          // cached lpost(x); -inf until first known (fit-numpy.py:66)
  double k0;            // HMC: kinetic energy at the start of the trajectory
  const double* z;      // replay normals [steps][p]
  const double* u;      // replay uniforms [steps]
  double* out;          // thinned samples [iters][p]
  double x[kMaxP], gx[kMaxP], q[kMaxP], mom[kMaxP];
  double scale[kMaxP], sqrt_scale[kMaxP];
  double beta_in[kMaxP];  // where the next fused evaluation happens
};

struct FinishArgs {
  const double* beta;        // the evaluated point (device, p doubles)
  const double* pscale;      // prior sd, p doubles
  const double* log_pscale;  // log(pscale)
  double* res;               // [lpost, ll, lprior, glp[0..p)]
  SamplerState* state;       // nullptr for a bare evaluation
  int32_t p, world, rank, p2p;
  // fused peer-memory allreduce (comm.cuh)
  double* mailbox_local;
  unsigned long long* flags_local;
  double* mailbox_peer[kMaxRanks];
  unsigned long long* flags_peer[kMaxRanks];
  unsigned long long* seq;
  int* comm_error;           // set to 1 if a peer's sums did not arrive within ~4 s (a rank died)
};

constexpr int kMailStride = kMaxP + 1;  // doubles per (slot, rank) mailbox entry

__device__ inline double draw_z(const SamplerState* st, long long t, int j) {
  if (st->rng == RNG_REPLAY) return st->z[(t - st->t_replay0) * st->p + j];
  return philox_normal(st->seed, (uint64_t)t, (uint32_t)j);
}
__device__ inline double draw_u(const SamplerState* st, long long t) {
  if (st->rng == RNG_REPLAY) return st->u[t - st->t_replay0];
  return philox_uniform(st->seed, (uint64_t)t);
}

// Generate the next evaluation point from the current state (x, gx) with the
// draws of iteration st->t.  Called by the whole CTA.
__device__ inline void sampler_propose(SamplerState* st, double* scratch) {
  const int j = threadIdx.x;
  const bool act = j < st->p;
  const long long t = st->t;
  const int kind = st->kind;
  double ksum = 0.0;
  if (act) {
    const double x = st->x[j];
    if (kind == S_RWMH) {
      // fit-numpy.py:84: beta + 0.02*pre*randn(p), scale = 0.02*pre
      st->beta_in[j] = x + st->scale[j] * draw_z(st, t, j);
    } else if (kind == S_UL) {
      st->beta_in[j] = x;  // the evaluation at x IS the step (fit-np-ul.py:65)
    } else if (kind == S_MALA) {
      // fit-np-mala.py:76-77: advance(x) + randn(p)*spre*sdt
      const double adv = x + 0.5 * st->scale[j] * st->gx[j] * st->step;
      const double prop = adv + draw_z(st, t, j) * st->sqrt_scale[j] * st->sqrt_step;
      st->q[j] = prop;
      st->beta_in[j] = prop;
    } else {
      // fit-np-hmc.py:85 p = randn(d)*sdmm; :68 half kick with the cached gradient; :70 drift
      const double dmm = st->scale[j];
      double m = draw_z(st, t, j) * st->sqrt_scale[j];
      ksum = (m * m) / dmm;
      m = m + 0.5 * st->step * st->gx[j];
      const double q = x + st->step * m / dmm;
      st->mom[j] = m;
      st->q[j] = q;
      st->beta_in[j] = q;
    }
  }
  if (kind == S_HMC) {
    const double k0 = 0.5 * block_sum(ksum, scratch);
    if (j == 0) { st->k0 = k0; st->leap = 0; }
  }
}

// Consume one fused evaluation (lp = lpost, g = this thread's glp component at
// st->beta_in) and advance the chain until the next evaluation point is known.
__device__ inline void sampler_on_eval(SamplerState* st, double lp, double g, double* scratch) {
  const int j = threadIdx.x;
  const int p = st->p;
  const bool act = j < p;
  const int kind = st->kind;
  const int phase = st->phase;
  const double lp_x = st->lp_x;
  const long long t = st->t;
  const int leap_now = st->leap;
  __syncthreads();  // everyone has read the scalars thread 0 is about to change
  if (phase == PH_PAUSED) return;

  if (phase == PH_INIT) {
    // evaluation at the initial state: cache its gradient (and, for HMC, lpost).
    // MALA/RWMH keep lp_x = -inf: the reference never evaluates lpost(init) and
    // always accepts the first proposal (fit-np-mala.py:82).
    if (act) st->gx[j] = g;
    if (j == 0) {
      if (kind == S_HMC) st->lp_x = lp;
      st->phase = PH_STEP;
    }
    __syncthreads();
    sampler_propose(st, scratch);
    return;
  }

  bool stepped = true;  // did this evaluation complete a kernel application?
  bool acc = false;

  if (kind == S_RWMH) {
    // fit-numpy.py:57-60 (dprop cancels): a = lp - ll; accept iff log(u) < a
    const double a = lp - lp_x;
    acc = log(draw_u(st, t)) < a;
    __syncthreads();
    if (acc && act) st->x[j] = st->beta_in[j];
  } else if (kind == S_UL) {
    // fit-np-ul.py:65-67
    if (act) {
      const double adv = st->x[j] + 0.5 * st->scale[j] * g * st->step;
      st->x[j] = adv + draw_z(st, t, j) * st->sqrt_scale[j] * st->sqrt_step;
    }
    acc = true;
  } else if (kind == S_MALA) {
    // fit-np-mala.py:65 with dprop(new, old) = sum logpdf(new; advance(old), spre*sdt):
    // a = lp - ll - 1/2 sum((x - adv(prop))/s)^2 + 1/2 sum((prop - adv(x))/s)^2
    double fwd = 0.0, bwd = 0.0;
    if (act) {
      const double pre = st->scale[j], s = st->sqrt_scale[j] * st->sqrt_step;
      const double x = st->x[j], prop = st->q[j];
      const double adv_x = x + 0.5 * pre * st->gx[j] * st->step;
      const double adv_p = prop + 0.5 * pre * g * st->step;
      const double zf = (prop - adv_x) / s, zb = (x - adv_p) / s;
      fwd = zf * zf;
      bwd = zb * zb;
    }
    fwd = block_sum(fwd, scratch);
    bwd = block_sum(bwd, scratch);
    const double a = lp - lp_x + (-0.5 * bwd) - (-0.5 * fwd);
    acc = log(draw_u(st, t)) < a;
    __syncthreads();
    if (acc && act) { st->x[j] = st->q[j]; st->gx[j] = g; }
  } else {
    const int leap = leap_now;
    if (leap < st->l - 1) {
      // fit-np-hmc.py:70-72: full kick then drift
      if (act) {
        const double m = st->mom[j] + st->step * g;
        const double q = st->q[j] + st->step * m / st->scale[j];
        st->mom[j] = m;
        st->q[j] = q;
        st->beta_in[j] = q;
      }
      __syncthreads();
      if (j == 0) st->leap = leap + 1;
      stepped = false;
    } else {
      // fit-np-hmc.py:74 last half kick; :59,:76-78 a = alpi(prop) - alpi(x)
      double ksum = 0.0;
      if (act) {
        const double m = st->mom[j] + 0.5 * st->step * g;
        st->mom[j] = m;
        ksum = (m * m) / st->scale[j];
      }
      const double k1 = 0.5 * block_sum(ksum, scratch);
      const double a = (lp - k1) - (lp_x - st->k0);
      acc = log(draw_u(st, t)) < a;
      __syncthreads();
      if (acc && act) { st->x[j] = st->q[j]; st->gx[j] = g; }
    }
  }
  if (!stepped) return;

  __syncthreads();
  const long long t1 = t + 1;
  const long long done_in_run = t1 - st->t_run0;
  const bool emit = (done_in_run % st->thin) == 0;
  const bool last = (t1 == st->t_end);
  if (emit && act) st->out[(done_in_run / st->thin - 1) * p + j] = st->x[j];
  __syncthreads();
  if (j == 0) {
    st->t = t1;
    if (acc) {
      if (kind != S_UL) st->lp_x = lp;
      st->accepted += 1;
    }
    if (last) st->phase = PH_PAUSED;
  }
  __syncthreads();
  if (!last) sampler_propose(st, scratch);
}

// Arm the state(s) for a run of `steps` kernel applications. init == nullptr
// continues from the paused chain.  Launched as <<<C, kBlock>>>: block c arms chain c
// (states[c], init + c*p, Philox key seed + c*golden, replay rows c*steps.., samples
// out + c*iters*p).
__global__ void sampler_begin_kernel(SamplerState* states, const double* init, const double* scale,
                                     int kind, int l, int p, int rng, double step, uint64_t seed,
                                     double init_lpost, long long steps, long long thin, const double* z,
                                     const double* u, double* out, int reuse_cache) {
  __shared__ double scratch[kWarps];
  const int j = threadIdx.x;
  const long long c = blockIdx.x;
  SamplerState* st = states + c;
  if (init) init += c * p;
  if (j < p) {
    st->scale[j] = scale[j];
    st->sqrt_scale[j] = sqrt(scale[j]);
    if (init) st->x[j] = init[j];
  }
  if (j == 0) {
    st->kind = kind; st->l = l; st->p = p; st->rng = rng;
    st->step = step; st->sqrt_step = sqrt(step);
    st->seed = seed + (uint64_t)c * 0x9E3779B97F4A7C15ull;
    if (init) {
      st->t = 0; st->accepted = 0; st->k0 = 0.0;
      if (!(reuse_cache && kind == S_HMC)) st->lp_x = init_lpost;   // HMC keeps the cached lpost(x)
    }
    st->t_run0 = st->t;
    st->t_replay0 = st->t;
    st->t_end = st->t + steps;
    st->thin = thin;
    st->z = z ? z + c * steps * p : nullptr;
    st->u = u ? u + c * steps : nullptr;
    st->out = out + c * (steps / thin) * p;
    st->leap = 0;
  }
  __syncthreads();
  if (steps <= 0) { if (j == 0) st->phase = PH_PAUSED; return; }
  // Which samplers need the gradient (and lpost) of the starting state first?
  // A continued MALA/HMC chain still has them cached.
  const bool need_init = init != nullptr && !reuse_cache && (kind == S_MALA || kind == S_HMC);
  if (need_init) {
    if (j < p) st->beta_in[j] = st->x[j];
    if (j == 0) st->phase = PH_INIT;
  } else {
    if (j == 0) st->phase = PH_STEP;
    __syncthreads();
    sampler_propose(st, scratch);
  }
}

// Warm L1 with everything finish_eval / sampler_on_eval will read, so those (dependent) loads
// overlap the reduction of the CTA partials instead of following it.  Call only after the
// gpu-scope fence that follows the ticket (it invalidates this SM's L1).
__device__ inline void prefetch_finish_inputs(const FinishArgs& f) {
  const int j = threadIdx.x;
  auto pf = [](const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); };
  if (j < f.p) {
    pf(f.beta + j); pf(f.pscale + j); pf(f.log_pscale + j);
    if (f.state) {
      const SamplerState* st = f.state;
      pf(st->x + j); pf(st->gx + j); pf(st->q + j); pf(st->mom + j); pf(st->scale + j); pf(st->sqrt_scale + j);
    }
  }
  if (j == 0 && f.state) pf(f.state);
}

// ---------------------------------------------------------------- finish
// sums (shared memory, p+1 doubles): [ll, X'(y-p)] summed over all rows of this
// rank.  Combines across ranks (fused peer-memory mode), adds the prior
// (fit-np-ul.py:33-34,46), publishes [lpost, ll, glp] and feeds the sampler.
__device__ inline void finish_eval(const FinishArgs& f, double* sums, double* scratch) {
  const int j = threadIdx.x;
  const int p = f.p;
  const bool act = j < p;

  if (f.p2p && f.world > 1) {
    // One-shot allreduce over NVLink peer memory: store my sums into slot
    // (seq&1, my rank) of every peer's mailbox, publish a flag, wait for every
    // peer's flag, add in rank order.  Two slots suffice: a peer cannot run two
    // evaluations ahead because it needs my sums of evaluation k to finish k.
    const unsigned long long seq = *f.seq;
    const int slot = (int)(seq & 1ull);
    for (int r = 0; r < f.world; ++r) {
      if (r == f.rank) continue;
      double* dst = f.mailbox_peer[r] + ((size_t)slot * kMaxRanks + f.rank) * kMailStride;
      for (int c = j; c <= p; c += kBlock) dst[c] = sums[c];
    }
    // No system fence per thread: the CTA barrier orders every thread's remote stores before the
    // st.release.sys of the flag (release is cumulative over what happens-before it), which is
    // the single system-scope fence of the exchange.
    if (f.p2p & 2) __threadfence_system();   // A/B knob (LRB_P2P_FENCE=1): the conservative variant
    __syncthreads();
    if (j < f.world && j != f.rank) {
      unsigned long long* fl = f.flags_peer[j] + (size_t)slot * kMaxRanks + f.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(seq + 1) : "memory");
      const unsigned long long* mine = f.flags_local + (size_t)slot * kMaxRanks + j;
      unsigned long long got;
      const long long t_start = clock64();
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
        if (got < seq + 1 && clock64() - t_start > 8000000000ll) {   // ~4 s: a peer is gone; do not hang the GPU
          if (f.comm_error) *f.comm_error = 1;
          break;
        }
      } while (got < seq + 1);
    }
    __syncthreads();
    for (int c = j; c <= p; c += kBlock) {
      double tot = 0.0;
      for (int r = 0; r < f.world; ++r) {
        const double* src = f.mailbox_local + ((size_t)slot * kMaxRanks + r) * kMailStride;
        tot += (r == f.rank) ? sums[c] : __ldcg(src + c);
      }
      sums[c] = tot;  // each thread reads and writes only its own c
    }
    __syncthreads();
    if (j == 0) *f.seq = seq + 1;
  }

  const double LOG_SQRT_2PI = 0.91893853320467274178;
  double lpr = 0.0, g = 0.0;
  if (act) {
    const double b = f.beta[j], ps = f.pscale[j];
    const double zz = b / ps;
    lpr = -(zz * zz) / 2.0 - LOG_SQRT_2PI - f.log_pscale[j];
    g = -b / (ps * ps) + sums[1 + j];
  }
  const double lprior = block_sum(lpr, scratch);
  const double ll = sums[0];
  const double lpost = ll + lprior;
  if (j == 0) { f.res[0] = lpost; f.res[1] = ll; f.res[2] = lprior; }
  if (act) f.res[3 + j] = g;
  if (f.state) {
    __syncthreads();
    sampler_on_eval(f.state, lpost, g, scratch);
  }
}

// lprior only (no data pass): out[c] = sum_j logpdf(beta[c][j]; 0, pscale[j])
__global__ void prior_kernel(const double* beta, const double* pscale, const double* log_pscale,
                             int p, double* out) {
  __shared__ double scratch[kWarps];
  const double LOG_SQRT_2PI = 0.91893853320467274178;
  const int j = threadIdx.x;
  double lpr = 0.0;
  if (j < p) {
    const double zz = beta[(size_t)blockIdx.x * p + j] / pscale[j];
    lpr = -(zz * zz) / 2.0 - LOG_SQRT_2PI - log_pscale[j];
  }
  const double s = block_sum(lpr, scratch);
  if (j == 0) out[blockIdx.x] = s;
}

// Separate finish launch (NCCL mode: runs after the allreduce of `sums`).
__global__ void finish_kernel(FinishArgs f, const double* sums_global) {
  __shared__ double sums[kMaxP + 1];
  __shared__ double scratch[kWarps];
  if (f.state && f.state->phase == PH_PAUSED) return;
  for (int c = threadIdx.x; c <= f.p; c += kBlock) sums[c] = sums_global[c];
  __syncthreads();
  finish_eval(f, sums, scratch);
}

}  // namespace lrb
