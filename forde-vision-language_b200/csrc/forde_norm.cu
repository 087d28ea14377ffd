// Fused residual-add + LayerNorm (the `add/norm` either side of the StatefulLayer in a FORDE block,
// README.md:42-43).  HBM-bound streaming kernels: one warp per token row, 16-byte vector accesses, the row
// stays in registers between the statistics and the normalisation, persistent grid (a multiple of the SM
// count).  fwd traffic = read x, r + write out; bwd = read dy, x, r + write ds; d(gamma), d(beta) are reduced
// per CTA in registers/smem and combined by a second tiny kernel in a fixed order (deterministic).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "../../include/forde_b200.h"
#include "forde_host.h"

namespace forde {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 4;  // 16-byte vectors per lane: cols <= 32 * 4 * 8 = 1024

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int NV>  // 16-byte vectors per lane actually used: ceil(cols / 256)
__global__ void __launch_bounds__(LN_WARPS * 32)
add_layernorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ r, const float* __restrict__ gamma,
                         const float* __restrict__ beta, uint4* __restrict__ out, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, int rows, int cols, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const float inv_n = 1.f / static_cast<float>(cols);
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const size_t base = static_cast<size_t>(row) * nvec;
    float s[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float a[8], b[8];
        unpack8(__ldg(x + base + v), a);
        if (r != nullptr) {
          unpack8(__ldg(r + base + v), b);
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] += b[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s[i][k] = a[k];
          sum += a[k];
        }
      }
    }
    const float mean = wsum(sum) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dlt = s[i][k] - mean;
          sq = fmaf(dlt, dlt, sq);
        }
      }
    }
    const float rstd = rsqrtf(wsum(sq) * inv_n + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf((s[i][k] - mean) * rstd, g[k], b[k]);
        out[base + v] = pack8(o);
      }
    }
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// partials: [gridDim.x][2][cols]  (d_gamma, d_beta)
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 3 ? 2 : 1)
add_layernorm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ r,
                         const float* __restrict__ gamma, const float* __restrict__ mean_in,
                         const float* __restrict__ rstd_in, uint4* __restrict__ ds, float* __restrict__ partials,
                         int rows, int cols) {
  extern __shared__ float red[];  // [LN_WARPS][2][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const float inv_n = 1.f / static_cast<float>(cols);
  float dg[NV][8], db[NV][8], g[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dg[i][k] = 0.f;
      db[i][k] = 0.f;
      g[i][k] = (v < nvec) ? __ldg(gamma + v * 8 + k) : 0.f;
    }
  }
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const size_t base = static_cast<size_t>(row) * nvec;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NV][8], gy[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float a[8], b[8], d[8];
        unpack8(__ldg(x + base + v), a);
        if (r != nullptr) {
          unpack8(__ldg(r + base + v), b);
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] += b[k];
        }
        unpack8(__ldg(dy + base + v), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float h = (a[k] - mean) * rstd;
          xh[i][k] = h;
          dg[i][k] = fmaf(d[k], h, dg[i][k]);
          db[i][k] += d[k];
          const float t = d[k] * g[i][k];
          gy[i][k] = t;
          c1 += t;
          c2 = fmaf(t, h, c2);
        }
      }
    }
    c1 = wsum(c1) * inv_n;
    c2 = wsum(c2) * inv_n;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rstd * (gy[i][k] - c1 - xh[i][k] * c2);
        ds[base + v] = pack8(o);
      }
    }
  }
  // CTA reduction of d_gamma / d_beta over the warps, fixed order
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        red[(warp * 2 + 0) * cols + v * 8 + k] = dg[i][k];
        red[(warp * 2 + 1) * cols + v * 8 + k] = db[i][k];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    const int which = c / cols, col = c - which * cols;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) t += red[(w * 2 + which) * cols + col];
    partials[(static_cast<size_t>(blockIdx.x) * 2 + which) * cols + col] = t;
  }
}

// d_gamma / d_beta from the per-CTA partials: block = 32 columns x 8 lanes over the CTAs, fixed combination order.
__global__ void ln_param_grad_kernel(const float* __restrict__ partials, int nblocks, int cols,
                                     float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;  // index into [2][cols]
  float t = 0.f;
  if (c < 2 * cols) {
    const int which = c / cols, col = c - which * cols;
    for (int b = threadIdx.y; b < nblocks; b += 8) t += partials[(static_cast<size_t>(b) * 2 + which) * cols + col];
  }
  red[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && c < 2 * cols) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) r += red[y][threadIdx.x];
    const int which = c / cols, col = c - which * cols;
    (which == 0 ? d_gamma : d_beta)[col] += r;
  }
}

static int ln_grid(int rows) {
  const int sms = sm_count();
  if (sms <= 0) return 0;
  int grid = 4 * sms;
  const int need = (rows + LN_WARPS - 1) / LN_WARPS;
  return grid < need ? grid : need;
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_ln_num_partials(int rows) { return rows <= 0 ? 0 : ln_grid(rows); }

int forde_add_layernorm_fwd(void* stream, const void* x, const void* r, const float* gamma, const float* beta,
                            void* out, float* mean, float* rstd, int rows, int cols, float eps) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || out == nullptr || mean == nullptr ||
      rstd == nullptr || rows <= 0 || cols <= 0)
    return FORDE_EINVAL;
  if ((cols & 7) != 0 || cols > 32 * LN_MAXV * 8 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(r) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) != 0 || (reinterpret_cast<uintptr_t>(beta) & 15) != 0)
    return FORDE_EALIGN;
  const int grid = ln_grid(rows);
  if (grid <= 0) return FORDE_ENODEV;
  const int nv = (cols / 8 + 31) / 32;
#define FORDE_LN_FWD(NV)                                                                                      \
  add_layernorm_fwd_kernel<NV><<<grid, LN_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(               \
      static_cast<const uint4*>(x), static_cast<const uint4*>(r), gamma, beta, static_cast<uint4*>(out), mean, \
      rstd, rows, cols, eps)
  if (nv == 1) FORDE_LN_FWD(1);
  else if (nv == 2) FORDE_LN_FWD(2);
  else if (nv == 3) FORDE_LN_FWD(3);
  else FORDE_LN_FWD(4);
#undef FORDE_LN_FWD
  return check_launch("add_layernorm_fwd_kernel");
}

int forde_add_layernorm_bwd(void* stream, const void* dy, const void* x, const void* r, const float* gamma,
                            const float* mean, const float* rstd, void* ds, float* partials, float* d_gamma,
                            float* d_beta, int rows, int cols) {
  if (dy == nullptr || x == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr || ds == nullptr ||
      partials == nullptr || d_gamma == nullptr || d_beta == nullptr || rows <= 0 || cols <= 0)
    return FORDE_EINVAL;
  if ((cols & 7) != 0 || cols > 32 * LN_MAXV * 8 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(r) & 15) != 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(ds) & 15) != 0)
    return FORDE_EALIGN;
  const int grid = ln_grid(rows);
  if (grid <= 0) return FORDE_ENODEV;
  const int smem = LN_WARPS * 2 * cols * static_cast<int>(sizeof(float));
  const int max_smem = LN_WARPS * 2 * 32 * LN_MAXV * 8 * static_cast<int>(sizeof(float));
  static std::once_flag once;
  std::call_once(once, [max_smem] {
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  });
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (cols / 8 + 31) / 32;
#define FORDE_LN_BWD(NV)                                                                                         \
  add_layernorm_bwd_kernel<NV><<<grid, LN_WARPS * 32, smem, st>>>(                                               \
      static_cast<const uint4*>(dy), static_cast<const uint4*>(x), static_cast<const uint4*>(r), gamma, mean, rstd, \
      static_cast<uint4*>(ds), partials, rows, cols)
  if (nv == 1) FORDE_LN_BWD(1);
  else if (nv == 2) FORDE_LN_BWD(2);
  else if (nv == 3) FORDE_LN_BWD(3);
  else FORDE_LN_BWD(4);
#undef FORDE_LN_BWD
  ln_param_grad_kernel<<<(2 * cols + 31) / 32, dim3(32, 8), 0, st>>>(partials, grid, cols, d_gamma, d_beta);
  return check_launch("add_layernorm_bwd_kernel", 2);
}

}  // extern "C"
