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
#ifndef FORDE_LN_BWD_BLOCKS
#define FORDE_LN_BWD_BLOCKS(NV) ((NV) <= 3 ? 3 : 2)
#endif
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
// Register diet (ncu: the first version held d_gamma / d_beta / gamma / x_hat / g*dy in registers, 128 regs, and
// serialised its loads -- 77 % of the stall samples were long-scoreboard waits): the row's three inputs stay PACKED
// (bf16) in registers so all nine 16-byte loads issue back to back, x_hat and gamma*dy are recomputed for the second
// pass, and the per-column d_gamma / d_beta accumulate in the warp's own slice of shared memory (each lane owns its
// columns, so no conflicts and no atomics).
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, FORDE_LN_BWD_BLOCKS(NV))
add_layernorm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ r,
                         const float* __restrict__ gamma, const float* __restrict__ mean_in,
                         const float* __restrict__ rstd_in, uint4* __restrict__ ds, float* __restrict__ partials,
                         int rows, int cols) {
  extern __shared__ float red[];  // [LN_WARPS][2][cols] running d_gamma / d_beta of each warp, then [cols] gamma
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const float inv_n = 1.f / static_cast<float>(cols);
  float* gam_s = red + LN_WARPS * 2 * cols;
  float* dg_s = red + (warp * 2 + 0) * cols;
  float* db_s = red + (warp * 2 + 1) * cols;
  for (int c = threadIdx.x; c < LN_WARPS * 2 * cols; c += blockDim.x) red[c] = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) gam_s[c] = gamma[c];
  __syncthreads();
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const size_t base = static_cast<size_t>(row) * nvec;
    uint4 px[NV], pr[NV], pd[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        px[i] = __ldg(x + base + v);
        pr[i] = (r != nullptr) ? __ldg(r + base + v) : make_uint4(0u, 0u, 0u, 0u);
        pd[i] = __ldg(dy + base + v);
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float a[8], b[8], d[8];
        unpack8(px[i], a);
        unpack8(pr[i], b);
        unpack8(pd[i], d);
        float4* dgp = reinterpret_cast<float4*>(dg_s + v * 8);
        float4* dbp = reinterpret_cast<float4*>(db_s + v * 8);
        const float4* gp = reinterpret_cast<const float4*>(gam_s + v * 8);
        float4 g0 = dgp[0], g1 = dgp[1], b0 = dbp[0], b1 = dbp[1];
        const float4 w0 = gp[0], w1 = gp[1];
        float dgv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float dbv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float gw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float h = (a[k] + b[k] - mean) * rstd;
          dgv[k] = fmaf(d[k], h, dgv[k]);
          dbv[k] += d[k];
          const float t = d[k] * gw[k];
          c1 += t;
          c2 = fmaf(t, h, c2);
        }
        dgp[0] = make_float4(dgv[0], dgv[1], dgv[2], dgv[3]);
        dgp[1] = make_float4(dgv[4], dgv[5], dgv[6], dgv[7]);
        dbp[0] = make_float4(dbv[0], dbv[1], dbv[2], dbv[3]);
        dbp[1] = make_float4(dbv[4], dbv[5], dbv[6], dbv[7]);
      }
    }
    c1 = wsum(c1) * inv_n;
    c2 = wsum(c2) * inv_n;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float a[8], b[8], d[8], o[8];
        unpack8(px[i], a);
        unpack8(pr[i], b);
        unpack8(pd[i], d);
        const float4* gp = reinterpret_cast<const float4*>(gam_s + v * 8);
        const float4 w0 = gp[0], w1 = gp[1];
        const float gw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float h = (a[k] + b[k] - mean) * rstd;
          o[k] = rstd * (d[k] * gw[k] - c1 - h * c2);
        }
        ds[base + v] = pack8(o);
      }
    }
  }
  // CTA reduction of d_gamma / d_beta over the warps, fixed order
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    const int which = c / cols, col = c - which * cols;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) t += red[(w * 2 + which) * cols + col];
    partials[(static_cast<size_t>(blockIdx.x) * 2 + which) * cols + col] = t;
  }
}

// d_gamma / d_beta from the per-CTA partials: block = 32 columns x 32 lanes over the CTAs, fixed combination order.
__global__ void ln_param_grad_kernel(const float* __restrict__ partials, int nblocks, int cols,
                                     float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;  // index into [2][cols]
  float t = 0.f;
  if (c < 2 * cols) {
    const int which = c / cols, col = c - which * cols;
    for (int b = threadIdx.y; b < nblocks; b += 32) t += partials[(static_cast<size_t>(b) * 2 + which) * cols + col];
  }
  red[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && c < 2 * cols) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) r += red[y][threadIdx.x];
    const int which = c / cols, col = c - which * cols;
    (which == 0 ? d_gamma : d_beta)[col] += r;
  }
}

// Persistent grids: exactly the number of CTAs that are resident at once (a grid-stride loop over a larger grid
// would run a second, mostly empty wave), asked from the occupancy calculator per kernel instance.
template <typename K>
static int resident_blocks_per_sm(K kernel, int smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, LN_WARPS * 32, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}
static int clamp_grid(int per_sm, int rows) {
  const int sms = sm_count();
  if (sms <= 0) return 0;
  const int grid = per_sm * sms;
  const int need = (rows + LN_WARPS - 1) / LN_WARPS;
  return grid < need ? grid : need;
}
static int bwd_smem(int cols) { return (LN_WARPS * 2 + 1) * cols * static_cast<int>(sizeof(float)); }
static void bwd_prepare() {
  static std::once_flag once;
  std::call_once(once, [] {
    const int max_smem = bwd_smem(32 * LN_MAXV * 8);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(add_layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  });
}
static int bwd_grid(int rows, int cols) {
  bwd_prepare();
  const int nv = (cols / 8 + 31) / 32, smem = bwd_smem(cols);
  const int per_sm = nv == 1   ? resident_blocks_per_sm(add_layernorm_bwd_kernel<1>, smem)
                     : nv == 2 ? resident_blocks_per_sm(add_layernorm_bwd_kernel<2>, smem)
                     : nv == 3 ? resident_blocks_per_sm(add_layernorm_bwd_kernel<3>, smem)
                               : resident_blocks_per_sm(add_layernorm_bwd_kernel<4>, smem);
  return clamp_grid(per_sm, rows);
}
static int fwd_grid(int rows, int cols) {
  const int nv = (cols / 8 + 31) / 32;
  const int per_sm = nv == 1   ? resident_blocks_per_sm(add_layernorm_fwd_kernel<1>, 0)
                     : nv == 2 ? resident_blocks_per_sm(add_layernorm_fwd_kernel<2>, 0)
                     : nv == 3 ? resident_blocks_per_sm(add_layernorm_fwd_kernel<3>, 0)
                               : resident_blocks_per_sm(add_layernorm_fwd_kernel<4>, 0);
  return clamp_grid(per_sm, rows);
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_ln_num_partials(int rows) {
  // upper bound over the supported widths: the narrowest rows allow the most resident CTAs
  return rows <= 0 ? 0 : bwd_grid(rows, 8);
}

int forde_add_layernorm_fwd(void* stream, const void* x, const void* r, const float* gamma, const float* beta,
                            void* out, float* mean, float* rstd, int rows, int cols, float eps) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || out == nullptr || mean == nullptr ||
      rstd == nullptr || rows <= 0 || cols <= 0)
    return FORDE_EINVAL;
  if ((cols & 7) != 0 || cols > 32 * LN_MAXV * 8 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(r) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) != 0 || (reinterpret_cast<uintptr_t>(beta) & 15) != 0)
    return FORDE_EALIGN;
  const int grid = fwd_grid(rows, cols);
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
  const int grid = bwd_grid(rows, cols);
  if (grid <= 0) return FORDE_ENODEV;
  const int smem = bwd_smem(cols);
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
  ln_param_grad_kernel<<<(2 * cols + 31) / 32, dim3(32, 32), 0, st>>>(partials, grid, cols, d_gamma, d_beta);
  return check_launch("add_layernorm_bwd_kernel", 2);
}

}  // extern "C"
