// data.cuh -- getting the script globals X, y into the HBM layout the fused kernel
// streams (row-major, zero-padded to P columns, mode dtype; y one byte per row),
// and the on-device synthetic problem of SURVEY.md section 8d.
//
// Reference side: Python/fit-numpy.py:12-19 builds X column-major float64 (pandas
// to_numpy + hstack) and y float32; here that conversion is a tiled transpose.
#pragma once
#include "common.cuh"

namespace lrb {

// dst[(r0+r)*P + c] = (T) src(r, c) for r < nr, c < P (zero for c >= p).
// src is row-major (ld = row stride) or column-major (ld = column stride).
// 32x32 tiles through shared memory so both sides stay coalesced.
template <typename ST, typename T, bool COLMAJOR>
__global__ void ingest_kernel(const ST* __restrict__ src, long long ld, long long nr, int p,
                              T* __restrict__ dst, int P) {
  __shared__ T tile[32][33];
  const long long r_base = (long long)blockIdx.x * 32;
  const int c_base = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (COLMAJOR) {
    // read: tx along rows (contiguous in a column-major source)
    for (int k = ty; k < 32; k += 8) {
      const long long r = r_base + tx;
      const int c = c_base + k;
      T v = (T)0;
      if (r < nr && c < p) v = (T)src[(long long)c * ld + r];
      tile[k][tx] = v;  // tile[col][row]
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const long long r = r_base + k;
      const int c = c_base + tx;
      if (r < nr && c < P) dst[r * P + c] = tile[tx][k];
    }
  } else {
    for (int k = ty; k < 32; k += 8) {
      const long long r = r_base + k;
      const int c = c_base + tx;
      if (r < nr && c < P) dst[r * P + c] = (c < p) ? (T)src[r * ld + c] : (T)0;
    }
  }
}

// y (float / double / u8, values in {0,1}) -> u8; *bad is set if any value is not 0 or 1.
template <typename YT>
__global__ void ingest_y_kernel(const YT* __restrict__ src, long long n, uint8_t* __restrict__ dst,
                                int* bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const YT v = src[i];
  const bool one = (v == (YT)1), zero = (v == (YT)0);
  if (!one && !zero) atomicExch(bad, 1);
  dst[i] = one ? 1 : 0;
}

// Synthetic X: column 0 is 1, columns 1..p-1 are N(0,1) float32 values (stored as T,
// so the FP32 and FP64 modes see the same numbers), columns >= p are 0.
// One thread per (row, 4-column group); Philox counter = (global row, group).
template <typename T>
__global__ void synth_x_kernel(T* __restrict__ X, long long n, int p, int P, uint64_t seed,
                               long long row_offset) {
  const int groups = P / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * groups) return;
  const long long r = idx / groups;
  const int gq = (int)(idx % groups);
  const unsigned long long gr = (unsigned long long)(r + row_offset);
  const Philox4 h = philox4x32_10((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)gq, kStreamDataX,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u1 = ((float)(h.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (float)h.y * k;
  const float u3 = ((float)(h.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u4 = (float)h.w * k;
  const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
  float z[4];
  z[0] = ra * cospif(2.0f * u2);
  z[1] = ra * sinpif(2.0f * u2);
  z[2] = rb * cospif(2.0f * u4);
  z[3] = rb * sinpif(2.0f * u4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = gq * 4 + i;
    float v = z[i];
    if (c == 0) v = 1.0f;
    if (c >= p) v = 0.0f;
    X[r * P + c] = (T)v;
  }
}

// y_i ~ Bernoulli(expit(x_i . beta_true)), uniform from Philox(global row).
template <typename T>
__global__ void synth_y_kernel(const T* __restrict__ X, long long n, int p, int P,
                               const double* __restrict__ beta_true, uint64_t seed,
                               long long row_offset, uint8_t* __restrict__ y) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double eta = 0.0;
  for (int c = 0; c < p; ++c) eta += (double)X[r * P + c] * beta_true[c];
  const unsigned long long gr = (unsigned long long)(r + row_offset);
  const Philox4 h = philox4x32_10((uint32_t)gr, (uint32_t)(gr >> 32), 0u, kStreamDataY,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = u01_53(h.x, h.y);
  y[r] = (u < 1.0 / (1.0 + exp(-eta))) ? 1 : 0;
}

// rows [row0, row0+nr) back out as float64 row-major (nr x p) + float32 y
template <typename T>
__global__ void export_rows_kernel(const T* __restrict__ X, const uint8_t* __restrict__ y,
                                   long long row0, long long nr, int p, int P,
                                   double* __restrict__ Xo, float* __restrict__ yo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nr * P) return;
  const long long r = idx / P;
  const int c = (int)(idx % P);
  if (c < p) Xo[r * p + c] = (double)X[(row0 + r) * P + c];
  if (c == 0) yo[r] = (float)y[row0 + r];
}

__global__ void rng_dump_kernel(uint64_t seed, long long t0, long long count, int p, double* z,
                                double* u) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count * p) {
    const long long t = t0 + idx / p;
    z[idx] = philox_normal(seed, (uint64_t)t, (uint32_t)(idx % p));
  }
  if (idx < count) u[idx] = philox_uniform(seed, (uint64_t)(t0 + idx));
}

}  // namespace lrb
