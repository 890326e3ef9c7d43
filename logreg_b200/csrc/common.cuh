// common.cuh -- shared device helpers: Philox4x32-10, block reductions, constants.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lrb {

constexpr int kBlock = 256;          // threads per CTA of every kernel in this library
constexpr int kWarps = kBlock / 32;
constexpr int kMaxP = 256;           // one thread per coefficient in the sampler update
constexpr int kMaxRanks = 8;         // one NVSwitch box

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al. 2011). The counter is (iteration, coordinate,
// stream) so every rank of a row-sharded run regenerates identical draws and a
// restarted chain continues the same stream (checkpoint row).
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// 53-bit uniform in the open interval (0,1) from two 32-bit words.
__host__ __device__ inline double u01_53(uint32_t hi, uint32_t lo) {
  uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);   // 27 + 26 bits
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

constexpr uint32_t kStreamNormal = 0u;   // sampler N(0,1) draws
constexpr uint32_t kStreamUniform = 1u;  // sampler accept uniforms
constexpr uint32_t kStreamDataX = 2u;    // synthetic design matrix
constexpr uint32_t kStreamDataY = 3u;    // synthetic responses
constexpr uint32_t kStreamSplit = 4u;    // key derivation of the keyed (JAX-style) front-end

// Child i of a 64-bit key: words (x, y) of Philox4x32-10 with counter (i_lo, i_hi, 0, kStreamSplit)
// under the parent key.  This is what logreg_b200.jaxlike.split(key, n)[i] returns
// (the role of jax.random.split in Python/fit-jax2.py:90,100,109).
__host__ __device__ inline uint64_t philox_child(uint64_t key, uint64_t i) {
  Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, kStreamSplit, (uint32_t)key, (uint32_t)(key >> 32));
  return ((uint64_t)r.y << 32) | (uint64_t)r.x;
}

// N(0,1) for (iteration t, coordinate j): Box-Muller on two 53-bit uniforms.
__device__ inline double philox_normal(uint64_t seed, uint64_t t, uint32_t j) {
  Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), j, kStreamNormal,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = u01_53(r.x, r.y), u2 = u01_53(r.z, r.w);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ inline double philox_uniform(uint64_t seed, uint64_t t) {
  Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0u, kStreamUniform,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01_53(r.x, r.y);
}

// ---------------------------------------------------------------- reductions
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA in a fixed order (deterministic; identical on every rank).
// `scratch` holds kWarps doubles. All threads receive the result.
__device__ inline double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) s += scratch[w];
  return s;
}

}  // namespace lrb
