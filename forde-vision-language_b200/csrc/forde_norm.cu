// (round 2) The arithmetic is written with Blackwell's packed fp32 pair instructions (FFMA2 / FADD2 / FMUL2 through
// __ffma2_rn / __fadd2_rn / __fmul2_rn): the kernels were issue-bound, not bandwidth-bound (the backward streamed
// 232 MB in 80 us = 2.9 TB/s), and a bf16 pair unpacks straight into a float2.
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
// a packed bf16 pair -> float2 {low element, high element}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
  __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int NV, bool HAS_R>  // NV: 16-byte vectors per lane actually used, ceil(cols / 256); HAS_R: r != NULL
__global__ void __launch_bounds__(LN_WARPS * 32)
add_layernorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ r, const float* __restrict__ gamma,
                         const float* __restrict__ beta, uint4* __restrict__ out, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, int rows, int cols, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const float inv_n = 1.f / static_cast<float>(cols);
  const int stride = gridDim.x * LN_WARPS;
  constexpr bool has_r = HAS_R;
  uint4 cx[NV], cr[NV];
  auto load_row = [&](int rw, uint4* ax, uint4* ar) {
    const size_t b0 = static_cast<size_t>(rw) * nvec;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        ax[i] = __ldg(x + b0 + v);
        if (has_r) ar[i] = __ldg(r + b0 + v);
      }
    }
  };
  int row = blockIdx.x * LN_WARPS + warp;
  if (row < rows) load_row(row, cx, cr);
  for (; row < rows; row += stride) {
    const size_t base = static_cast<size_t>(row) * nvec;
    uint4 nx[NV], nr[NV];
    if (row + stride < rows) load_row(row + stride, nx, nr);   // the next row's loads fly while this one is normalised
    float2 s[NV][4];
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint4 qx = cx[i];
        const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w};
        if (has_r) {
          const uint4 qr = cr[i];
          const uint32_t wr[4] = {qr.x, qr.y, qr.z, qr.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) s[i][k] = __fadd2_rn(unpack2(wx[k]), unpack2(wr[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) s[i][k] = unpack2(wx[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sum2 = __fadd2_rn(sum2, s[i][k]);
      }
    }
    const float mean = wsum(sum2.x + sum2.y) * inv_n;
    const float2 nmean2 = splat2(-mean);
    float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 dlt = __fadd2_rn(s[i][k], nmean2);
          sq2 = __ffma2_rn(dlt, dlt, sq2);
        }
      }
    }
    const float rstd = rsqrtf(wsum(sq2.x + sq2.y) * inv_n + eps);
    const float2 rstd2 = splat2(rstd), nmr2 = splat2(-mean * rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + 1);
        const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
        const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = pack2(__ffma2_rn(__ffma2_rn(s[i][k], rstd2, nmr2), g[k], b[k]));
        out[base + v] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx[i] = nx[i];
      if (has_r) cr[i] = nr[i];
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
    const float2 rstd2 = splat2(rstd), nmr2 = splat2(-mean * rstd);
    const bool has_r = (r != nullptr);   // uniform
    float2 c1_2 = make_float2(0.f, 0.f), c2_2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint32_t wx[4] = {px[i].x, px[i].y, px[i].z, px[i].w};
        const uint32_t wr[4] = {pr[i].x, pr[i].y, pr[i].z, pr[i].w};
        const uint32_t wd[4] = {pd[i].x, pd[i].y, pd[i].z, pd[i].w};
        float2* dgp = reinterpret_cast<float2*>(dg_s + v * 8);
        float2* dbp = reinterpret_cast<float2*>(db_s + v * 8);
        const float2* gp = reinterpret_cast<const float2*>(gam_s + v * 8);
        float4 g01 = reinterpret_cast<float4*>(dgp)[0], g23 = reinterpret_cast<float4*>(dgp)[1];
        float4 b01 = reinterpret_cast<float4*>(dbp)[0], b23 = reinterpret_cast<float4*>(dbp)[1];
        const float4 w01 = reinterpret_cast<const float4*>(gp)[0], w23 = reinterpret_cast<const float4*>(gp)[1];
        float2 dgv[4] = {make_float2(g01.x, g01.y), make_float2(g01.z, g01.w), make_float2(g23.x, g23.y), make_float2(g23.z, g23.w)};
        float2 dbv[4] = {make_float2(b01.x, b01.y), make_float2(b01.z, b01.w), make_float2(b23.x, b23.y), make_float2(b23.z, b23.w)};
        const float2 gw[4] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 a = unpack2(wx[k]);
          if (has_r) a = __fadd2_rn(a, unpack2(wr[k]));
          const float2 d = unpack2(wd[k]);
          const float2 h = __ffma2_rn(a, rstd2, nmr2);      // (x + r - mean) * rstd
          dgv[k] = __ffma2_rn(d, h, dgv[k]);
          dbv[k] = __fadd2_rn(dbv[k], d);
          const float2 t = __fmul2_rn(d, gw[k]);
          c1_2 = __fadd2_rn(c1_2, t);
          c2_2 = __ffma2_rn(t, h, c2_2);
        }
        reinterpret_cast<float4*>(dgp)[0] = make_float4(dgv[0].x, dgv[0].y, dgv[1].x, dgv[1].y);
        reinterpret_cast<float4*>(dgp)[1] = make_float4(dgv[2].x, dgv[2].y, dgv[3].x, dgv[3].y);
        reinterpret_cast<float4*>(dbp)[0] = make_float4(dbv[0].x, dbv[0].y, dbv[1].x, dbv[1].y);
        reinterpret_cast<float4*>(dbp)[1] = make_float4(dbv[2].x, dbv[2].y, dbv[3].x, dbv[3].y);
      }
    }
    const float c1 = wsum(c1_2.x + c1_2.y) * inv_n;
    const float c2 = wsum(c2_2.x + c2_2.y) * inv_n;
    const float2 nc2_2 = splat2(-c2), nc1r2 = splat2(-c1 * rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint32_t wx[4] = {px[i].x, px[i].y, px[i].z, px[i].w};
        const uint32_t wr[4] = {pr[i].x, pr[i].y, pr[i].z, pr[i].w};
        const uint32_t wd[4] = {pd[i].x, pd[i].y, pd[i].z, pd[i].w};
        const float4 w01 = reinterpret_cast<const float4*>(gam_s + v * 8)[0];
        const float4 w23 = reinterpret_cast<const float4*>(gam_s + v * 8)[1];
        const float2 gw[4] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 a = unpack2(wx[k]);
          if (has_r) a = __fadd2_rn(a, unpack2(wr[k]));
          const float2 h = __ffma2_rn(a, rstd2, nmr2);
          const float2 t = __fmul2_rn(unpack2(wd[k]), gw[k]);
          // rstd * (t - c1 - h * c2)
          o[k] = pack2(__ffma2_rn(__ffma2_rn(h, nc2_2, t), rstd2, nc1r2));
        }
        ds[base + v] = make_uint4(o[0], o[1], o[2], o[3]);
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

// Round-2 variant of the backward: per-lane REGISTER accumulators for d_gamma / d_beta (each lane owns its columns;
// flushed to shared memory once, at the end) and the next row's loads issued before the current row is processed.
// ncu of the shared-memory-accumulator version above: short-scoreboard (LDS -> FMA -> STS chains on the accumulators)
// and MIO-throttle stalls dominate, 232 MB streamed at 3.0 TB/s.
template <int NV, bool HAS_R>
__global__ void __launch_bounds__(LN_WARPS * 32, HAS_R ? 1 : 2)
add_layernorm_bwd_reg_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ r,
                             const float* __restrict__ gamma, const float* __restrict__ mean_in,
                             const float* __restrict__ rstd_in, uint4* __restrict__ ds, float* __restrict__ partials,
                             int rows, int cols) {
  extern __shared__ float red[];  // [LN_WARPS][2][cols] (used once, at the end), then [cols] gamma
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const float inv_n = 1.f / static_cast<float>(cols);
  float* gam_s = red + LN_WARPS * 2 * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) gam_s[c] = gamma[c];
  __syncthreads();
  float2 dg[NV][4], db[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) dg[i][k] = db[i][k] = make_float2(0.f, 0.f);

  const int stride = gridDim.x * LN_WARPS;
  int row = blockIdx.x * LN_WARPS + warp;
  uint4 px[NV], pr[NV], pd[NV];
  auto load_row = [&](int rw, uint4* ax, uint4* ar, uint4* ad) {
    const size_t base = static_cast<size_t>(rw) * nvec;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        ax[i] = __ldg(x + base + v);
        if (HAS_R) ar[i] = __ldg(r + base + v);
        ad[i] = __ldg(dy + base + v);
      }
    }
  };
  if (row < rows) load_row(row, px, pr, pd);
  while (row < rows) {
    const int nrow = row + stride;
    uint4 nx[NV], nr[NV], nd[NV];
    if (nrow < rows) load_row(nrow, nx, nr, nd);   // in flight while this row is processed
    const size_t base = static_cast<size_t>(row) * nvec;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float2 rstd2 = splat2(rstd), nmr2 = splat2(-mean * rstd);
    float2 c1_2 = make_float2(0.f, 0.f), c2_2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint32_t wx[4] = {px[i].x, px[i].y, px[i].z, px[i].w};
        const uint32_t wr[4] = {pr[i].x, pr[i].y, pr[i].z, pr[i].w};
        const uint32_t wd[4] = {pd[i].x, pd[i].y, pd[i].z, pd[i].w};
        const float4 w01 = reinterpret_cast<const float4*>(gam_s + v * 8)[0];
        const float4 w23 = reinterpret_cast<const float4*>(gam_s + v * 8)[1];
        const float2 gw[4] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 a = unpack2(wx[k]);
          if (HAS_R) a = __fadd2_rn(a, unpack2(wr[k]));
          const float2 d = unpack2(wd[k]);
          const float2 h = __ffma2_rn(a, rstd2, nmr2);      // (x + r - mean) * rstd
          dg[i][k] = __ffma2_rn(d, h, dg[i][k]);
          db[i][k] = __fadd2_rn(db[i][k], d);
          const float2 t = __fmul2_rn(d, gw[k]);
          c1_2 = __fadd2_rn(c1_2, t);
          c2_2 = __ffma2_rn(t, h, c2_2);
        }
      }
    }
    const float c1 = wsum(c1_2.x + c1_2.y) * inv_n;
    const float c2 = wsum(c2_2.x + c2_2.y) * inv_n;
    const float2 nc2_2 = splat2(-c2), nc1r2 = splat2(-c1 * rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint32_t wx[4] = {px[i].x, px[i].y, px[i].z, px[i].w};
        const uint32_t wr[4] = {pr[i].x, pr[i].y, pr[i].z, pr[i].w};
        const uint32_t wd[4] = {pd[i].x, pd[i].y, pd[i].z, pd[i].w};
        const float4 w01 = reinterpret_cast<const float4*>(gam_s + v * 8)[0];
        const float4 w23 = reinterpret_cast<const float4*>(gam_s + v * 8)[1];
        const float2 gw[4] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 a = unpack2(wx[k]);
          if (HAS_R) a = __fadd2_rn(a, unpack2(wr[k]));
          const float2 h = __ffma2_rn(a, rstd2, nmr2);
          const float2 t = __fmul2_rn(unpack2(wd[k]), gw[k]);
          o[k] = pack2(__ffma2_rn(__ffma2_rn(h, nc2_2, t), rstd2, nc1r2));   // rstd * (t - c1 - h * c2)
        }
        ds[base + v] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      px[i] = nx[i];
      if (HAS_R) pr[i] = nr[i];
      pd[i] = nd[i];
    }
    row = nrow;
  }
  // flush the lane-owned accumulators, then the fixed-order CTA reduction over the warps
  float* dg_s = red + (warp * 2 + 0) * cols;
  float* db_s = red + (warp * 2 + 1) * cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      reinterpret_cast<float4*>(dg_s + v * 8)[0] = make_float4(dg[i][0].x, dg[i][0].y, dg[i][1].x, dg[i][1].y);
      reinterpret_cast<float4*>(dg_s + v * 8)[1] = make_float4(dg[i][2].x, dg[i][2].y, dg[i][3].x, dg[i][3].y);
      reinterpret_cast<float4*>(db_s + v * 8)[0] = make_float4(db[i][0].x, db[i][0].y, db[i][1].x, db[i][1].y);
      reinterpret_cast<float4*>(db_s + v * 8)[1] = make_float4(db[i][2].x, db[i][2].y, db[i][3].x, db[i][3].y);
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
template <int NV, bool HAS_R>
static int launch_bwd_reg(cudaStream_t st, const void* dy, const void* x, const void* r, const float* gamma,
                          const float* mean, const float* rstd, void* ds, float* partials, int rows, int cols,
                          int* grid_out) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(add_layernorm_bwd_reg_kernel<NV, HAS_R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         bwd_smem(32 * LN_MAXV * 8));
  });
  const int smem = bwd_smem(cols);
  const int grid = clamp_grid(resident_blocks_per_sm(add_layernorm_bwd_reg_kernel<NV, HAS_R>, smem), rows);
  if (grid <= 0) return FORDE_ENODEV;
  add_layernorm_bwd_reg_kernel<NV, HAS_R><<<grid, LN_WARPS * 32, smem, st>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(x), static_cast<const uint4*>(r), gamma, mean, rstd,
      static_cast<uint4*>(ds), partials, rows, cols);
  *grid_out = grid;
  return FORDE_OK;
}

static int fwd_grid(int rows, int cols, bool has_r) {
  const int nv = (cols / 8 + 31) / 32;
#define FORDE_LN_OCC(NV) \
  (has_r ? resident_blocks_per_sm(add_layernorm_fwd_kernel<NV, true>, 0) : resident_blocks_per_sm(add_layernorm_fwd_kernel<NV, false>, 0))
  const int per_sm = nv == 1 ? FORDE_LN_OCC(1) : nv == 2 ? FORDE_LN_OCC(2) : nv == 3 ? FORDE_LN_OCC(3) : FORDE_LN_OCC(4);
#undef FORDE_LN_OCC
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
  const int grid = fwd_grid(rows, cols, r != nullptr);
  if (grid <= 0) return FORDE_ENODEV;
  const int nv = (cols / 8 + 31) / 32;
#define FORDE_LN_FWD(NV)                                                                                        \
  do {                                                                                                          \
    if (r != nullptr)                                                                                           \
      add_layernorm_fwd_kernel<NV, true><<<grid, LN_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(        \
          static_cast<const uint4*>(x), static_cast<const uint4*>(r), gamma, beta, static_cast<uint4*>(out),    \
          mean, rstd, rows, cols, eps);                                                                         \
    else                                                                                                        \
      add_layernorm_fwd_kernel<NV, false><<<grid, LN_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(       \
          static_cast<const uint4*>(x), static_cast<const uint4*>(r), gamma, beta, static_cast<uint4*>(out),    \
          mean, rstd, rows, cols, eps);                                                                         \
  } while (0)
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
  int grid = bwd_grid(rows, cols);
  if (grid <= 0) return FORDE_ENODEV;
  const int smem = bwd_smem(cols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (cols / 8 + 31) / 32;
  // Input already summed (r == NULL, what the FORDE block passes since the skip connection is added in the producing
  // GEMM's epilogue): register accumulators + next-row prefetch, 86 -> 66 us at 50 432 x 768.  With a separate
  // residual the extra row registers push that variant to one CTA per SM (105 us), so the shared-memory-accumulator
  // kernel stays for it (90 us).  profiles/r2_ln_variants.txt
  if (r == nullptr && nv <= 3) {   // (4 vectors per lane spill at two CTAs per SM: 147 us vs ~138 us at 65 792 x 1024)
    int rc = FORDE_OK;
#define FORDE_LN_BWD_REG(NV) \
  rc = launch_bwd_reg<NV, false>(st, dy, x, r, gamma, mean, rstd, ds, partials, rows, cols, &grid)
    if (nv == 1) FORDE_LN_BWD_REG(1);
    else if (nv == 2) FORDE_LN_BWD_REG(2);
    else FORDE_LN_BWD_REG(3);
#undef FORDE_LN_BWD_REG
    if (rc != FORDE_OK) return rc;
    ln_param_grad_kernel<<<(2 * cols + 31) / 32, dim3(32, 32), 0, st>>>(partials, grid, cols, d_gamma, d_beta);
    return check_launch("add_layernorm_bwd_reg_kernel", 2);
  }
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
