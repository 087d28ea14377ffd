// HBM-bound helper kernels of the FORDE hot path: sensing finalisation, bias-gradient column sums and the
// row/column log-softmax side of the CLIP contrastive loss.  All are simple coalesced streaming kernels; the
// heavy lifting (GEMMs + fused epilogues) lives in forde_gemm.cuh.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "../../include/forde_b200.h"
#include "forde_host.h"

namespace forde {

// ------------------------------------------------------------------ sensing (sensing.py:5-69)
// Hoyer sparsity of a length-n vector from its L1 norm and squared L2 norm, with the reference's guards:
// returns 0 for the all-zero vector (sensing.py:22,30) and for n == 1 (sensing.py:25,32).  fp32 like the
// reference (sensing.py:53-54).
__device__ __forceinline__ float hoyer_from_norms(float l1, float l2sq, int n) {
  if (n == 1) return 0.f;
  const float l2 = sqrtf(l2sq);
  if (l2 == 0.f) return 0.f;
  const float sn = sqrtf(static_cast<float>(n));
  return (sn - l1 / l2) / (sn - 1.f);
}

// stats row order (sensing.py:65-67): [grad_gini, grad_gdp, act_gini, act_gdp, act_variance]
__device__ __forceinline__ void stats_from_sums(double a_abs, double a_sum, double a_sq, double g_abs, double g_sq,
                                                int n, float* out5) {
  const double inv_n = 1.0 / static_cast<double>(n);
  const double mean = a_sum * inv_n;
  double var = a_sq * inv_n - mean * mean;  // jnp.var, ddof = 0 (sensing.py:59)
  if (var < 0.0) var = 0.0;
  out5[0] = hoyer_from_norms(static_cast<float>(g_abs), static_cast<float>(g_sq), n);
  out5[1] = static_cast<float>(g_abs * inv_n);
  out5[2] = hoyer_from_norms(static_cast<float>(a_abs), static_cast<float>(a_sq), n);
  out5[3] = static_cast<float>(a_abs * inv_n);
  out5[4] = static_cast<float>(var);
}

// block = FIN_NEUR neurons x FIN_LANES partial-lanes.  Lane y sums partial blocks y, y + FIN_LANES, ... in double,
// then the lanes are combined in a fixed order: the result does not depend on scheduling (bit-reproducible).
constexpr int FIN_LANES = 32;
constexpr int FIN_NEUR = 16;

__global__ void __launch_bounds__(FIN_NEUR * FIN_LANES)
stats_finalize_kernel(const float* __restrict__ act_p, const float* __restrict__ grad_p, int P, int H, int M,
                      float* __restrict__ step_stats, float* __restrict__ neuron_stats,
                      float* __restrict__ moments, float* __restrict__ moments_lo, int* sample_count,
                      int* step_count, float* __restrict__ db_up) {
  __shared__ double red[FIN_LANES][6][FIN_NEUR + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int h = blockIdx.x * FIN_NEUR + tx;
  if (blockIdx.x == 0 && tx == 0 && ty == 0) {
    if (sample_count != nullptr) *sample_count += M;
    if (step_count != nullptr) *step_count += 1;
  }
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (h < H) {
    for (int p = ty; p < P; p += FIN_LANES) {
      const size_t base = static_cast<size_t>(p) * 3 * H + h;
      if (act_p != nullptr) {
        acc[0] += act_p[base];
        acc[1] += act_p[base + H];
        acc[2] += act_p[base + 2 * static_cast<size_t>(H)];
      }
      if (grad_p != nullptr) {
        acc[3] += grad_p[base];
        acc[4] += grad_p[base + H];
        acc[5] += grad_p[base + 2 * static_cast<size_t>(H)];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) red[ty][k][tx] = acc[k];
  __syncthreads();
  if (ty != 0 || h >= H) return;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double t = 0;
#pragma unroll
    for (int y = 0; y < FIN_LANES; ++y) t += red[y][k][tx];
    acc[k] = t;
  }
  const double a_abs = acc[0], a_sum = acc[1], a_sq = acc[2], g_abs = acc[3], g_sum = acc[4], g_sq = acc[5];
  if (db_up != nullptr) db_up[h] += static_cast<float>(g_sum);
  float s[5];
  stats_from_sums(a_abs, a_sum, a_sq, g_abs, g_sq, M, s);
  if (step_stats != nullptr) {
#pragma unroll
    for (int k = 0; k < 5; ++k) step_stats[h * 5 + k] = s[k];
  }
  if (neuron_stats != nullptr) {
#pragma unroll
    for (int k = 0; k < 5; ++k) neuron_stats[h * 5 + k] += s[k];
  }
  if (moments != nullptr) {
    // The pooled variance is sum z^2 / n - mean^2 over a whole slow-loop interval (100 steps x 50k samples x ranks):
    // an fp32 running sum would lose it to cancellation for bias-dominated neurons (|mean| >> std).  The accumulator
    // is therefore a double-single pair of fp32 leaves (moments = high part, moments_lo = the rounding residue,
    // ~48 significant bits together); both stay ordinary fp32 arrays for the Flax-style collection.
    const double inc[5] = {a_abs, a_sum, a_sq, g_abs, g_sq};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double t = static_cast<double>(moments[h * 5 + k]) +
                       (moments_lo != nullptr ? static_cast<double>(moments_lo[h * 5 + k]) : 0.0) + inc[k];
      const float hi = static_cast<float>(t);
      moments[h * 5 + k] = hi;
      if (moments_lo != nullptr) moments_lo[h * 5 + k] = static_cast<float>(t - static_cast<double>(hi));
    }
  }
}

__global__ void stats_from_moments_kernel(const double* __restrict__ moments, const int* __restrict__ sample_count,
                                          int H, float* __restrict__ out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int n = *sample_count;
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (n > 0) {
    stats_from_sums(moments[h * 5 + 0], moments[h * 5 + 1], moments[h * 5 + 2], moments[h * 5 + 3],
                    moments[h * 5 + 4], n, s);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) out[h * 5 + k] = s[k];
}

// ------------------------------------------------------------------ column sums of a bf16 matrix (db_down)
// block = 32 x 8 threads: a warp covers 256 columns with 16-byte loads, 8 row lanes; grid.y splits the rows.
// (the scalar tail handles column counts that are not a multiple of 8 or unaligned rows)
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long lda, float* __restrict__ out,
                                   int rows, int cols, int rows_per_block, int vec_ok) {
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (vec_ok && c + 7 < cols) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(a + static_cast<size_t>(r) * lda + c));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[2 * i] += __uint_as_float(w[i] << 16);
        s[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  } else {
    for (int r = r0 + ty; r < r1; r += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c + i < cols) s[i] += __bfloat162float(a[static_cast<size_t>(r) * lda + c + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = s[i];
  __syncthreads();
  const int t = ty * 32 + tx;  // 256 threads: one column each
  float acc = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) acc += red[y][t];
  if (blockIdx.x * 256 + t < cols) atomicAdd(out + blockIdx.x * 256 + t, acc);
}

// ------------------------------------------------------------------ CLIP loss helpers (README.md:35)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per row: u = x / |x| (bf16), inv_norm = 1 / |x|
__global__ void l2norm_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ u,
                                   float* __restrict__ inv_norm, int B, int E, const float* __restrict__ log_temp,
                                   float* __restrict__ logit_scale_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (logit_scale_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *logit_scale_out = expf(*log_temp);
  if (row >= B) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * E;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float v = __bfloat162float(xr[e]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float rn = ss > 0.f ? rsqrtf(ss) : 0.f;
  for (int e = lane; e < E; e += 32) u[static_cast<size_t>(row) * E + e] = __float2bfloat16(__bfloat162float(xr[e]) * rn);
  if (lane == 0) inv_norm[row] = rn;
}

// one warp per row of S: row log-sum-exp and the diagonal (matching pair) logit
__global__ void row_lse_kernel(const float* __restrict__ S, int Bl, int Bg, int row_offset,
                               float* __restrict__ row_lse, float* __restrict__ diag) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Bl) return;
  const float* sr = S + static_cast<size_t>(row) * Bg;
  float m = -INFINITY;
  for (int j = lane; j < Bg; j += 32) m = fmaxf(m, sr[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < Bg; j += 32) s += __expf(sr[j] - m);
  s = warp_sum(s);
  if (lane == 0) {
    row_lse[row] = m + logf(s);
    diag[row] = sr[row_offset + row];
  }
}

// column max / sum-exp over the local rows.  block = 32 columns x 8 row lanes.
__global__ void col_lse_kernel(const float* __restrict__ S, int Bl, int Bg, float* __restrict__ col_max,
                               float* __restrict__ col_sum) {
  __shared__ float smax[8][32];
  __shared__ float ssum[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int j = blockIdx.x * 32 + tx;
  float m = -INFINITY, s = 0.f;
  if (j < Bg) {
    for (int i = ty; i < Bl; i += 8) {  // online softmax over the rows
      const float v = S[static_cast<size_t>(i) * Bg + j];
      const float nm = fmaxf(m, v);
      s = s * __expf(m - nm) + __expf(v - nm);
      m = nm;
    }
  }
  smax[ty][tx] = m;
  ssum[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < Bg) {
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < 8; ++y) M = fmaxf(M, smax[y][tx]);
    float T = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) T += (smax[y][tx] == -INFINITY) ? 0.f : ssum[y][tx] * __expf(smax[y][tx] - M);
    col_max[j] = M;
    col_sum[j] = T;
  }
}

// rescale a rank's column sum-exp to the global column max (before the cross-rank SUM all-reduce)
__global__ void clip_col_rescale_kernel(const float* __restrict__ col_max_local, const float* __restrict__ col_max_global,
                                        float* __restrict__ col_sum, int Bg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Bg) {
    const float ml = col_max_local[j];
    col_sum[j] = (ml == -INFINITY) ? 0.f : col_sum[j] * __expf(ml - col_max_global[j]);
  }
}

// col_lse = col_max + log(col_sum); loss += (sum_i (row_lse_i - diag_i) + sum_{j local} (col_lse_j - diag_j)) / (2 Bg)
// single block; B is at most a few thousand.
__global__ void clip_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ diag,
                                 const float* __restrict__ col_max, const float* __restrict__ col_sum, int Bl, int Bg,
                                 int row_offset, float* __restrict__ col_lse, float* __restrict__ loss) {
  float acc = 0.f;
  for (int j = threadIdx.x; j < Bg; j += blockDim.x) {
    const float l = col_max[j] + logf(col_sum[j]);
    col_lse[j] = l;
    const int i = j - row_offset;
    if (i >= 0 && i < Bl) acc += (l - diag[i]) + (row_lse[i] - diag[i]);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss, v * 0.5f / static_cast<float>(Bg));
  }
}

// dS = (softmax_row + softmax_col) / (2 Bg) - onehot / Bg  -> bf16 ; d(log_temp) partial += sum dS * S
__global__ void clip_ds_kernel(const float* __restrict__ S, const float* __restrict__ row_lse,
                               const float* __restrict__ col_lse, int Bl, int Bg, int row_offset,
                               __nv_bfloat16* __restrict__ dS, int ld_ds, float* __restrict__ d_scale) {
  const size_t total = static_cast<size_t>(Bl) * Bg;
  const float w = 0.5f / static_cast<float>(Bg);
  float acc = 0.f;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx / Bg), j = static_cast<int>(idx % Bg);
    const float s = S[idx];
    float g = (__expf(s - row_lse[i]) + __expf(s - col_lse[j])) * w;
    if (j == row_offset + i) g -= 2.f * w;
    dS[static_cast<size_t>(i) * ld_ds + j] = __float2bfloat16(g);
    acc = fmaf(g, s, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0 && d_scale != nullptr) atomicAdd(d_scale, v);
  }
}

// d_x = inv_norm * (dU - u (u . dU)); one warp per row
__global__ void l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ u, const float* __restrict__ inv_norm,
                                  const float* __restrict__ dU, float* __restrict__ dx, int B, int E) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const __nv_bfloat16* ur = u + static_cast<size_t>(row) * E;
  const float* dr = dU + static_cast<size_t>(row) * E;
  float dot = 0.f;
  for (int e = lane; e < E; e += 32) dot = fmaf(__bfloat162float(ur[e]), dr[e], dot);
  dot = warp_sum(dot);
  const float rn = inv_norm[row];
  for (int e = lane; e < E; e += 32)
    dx[static_cast<size_t>(row) * E + e] = rn * (dr[e] - __bfloat162float(ur[e]) * dot);
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_stats_finalize(void* stream, const float* act_partials, const float* grad_partials, int num_partials,
                         int H, int M, float* step_stats, float* neuron_stats, float* moments, float* moments_lo,
                         int* sample_count, int* step_count, float* db_up) {
  if (H <= 0 || M <= 0 || num_partials <= 0) return FORDE_EINVAL;
  if (act_partials == nullptr && grad_partials == nullptr) return FORDE_EINVAL;
  stats_finalize_kernel<<<(H + FIN_NEUR - 1) / FIN_NEUR, dim3(FIN_NEUR, FIN_LANES), 0, static_cast<cudaStream_t>(stream)>>>(
      act_partials, grad_partials, num_partials, H, M, step_stats, neuron_stats, moments, moments_lo, sample_count,
      step_count, db_up);
  return check_launch("stats_finalize_kernel");
}

int forde_stats_from_moments(void* stream, const double* moments, const int* sample_count, int H,
                             float* stats_out) {
  if (moments == nullptr || sample_count == nullptr || stats_out == nullptr || H <= 0) return FORDE_EINVAL;
  const int threads = 128;
  stats_from_moments_kernel<<<(H + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      moments, sample_count, H, stats_out);
  return check_launch("stats_from_moments_kernel");
}

int forde_colsum_bf16(void* stream, const void* a, long long lda, float* out, int rows, int cols) {
  if (a == nullptr || out == nullptr || rows <= 0 || cols <= 0) return FORDE_EINVAL;
  if ((lda & 1) != 0 || (reinterpret_cast<uintptr_t>(a) & 3) != 0) return FORDE_EALIGN;
  const int col_blocks = (cols + 255) / 256;
  int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  int row_blocks = (4 * sms + col_blocks - 1) / col_blocks;
  const int max_row_blocks = (rows + 63) / 64;
  if (row_blocks > max_row_blocks) row_blocks = max_row_blocks;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (rows + row_blocks - 1) / row_blocks;
  row_blocks = (rows + rows_per_block - 1) / rows_per_block;
  const int vec_ok = ((lda & 7) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) ? 1 : 0;
  colsum_bf16_kernel<<<dim3(col_blocks, row_blocks), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), lda, out, rows, cols, rows_per_block, vec_ok);
  return check_launch("colsum_bf16_kernel");
}

int forde_clip_fwd(void* stream, const void* img, const void* txt, const float* log_temp, int Bl, int Bg, int E,
                   int row_offset, void* u, void* v, float* inv_norm_img, float* inv_norm_txt,
                   float* logit_scale_out, float* S, float* row_lse, float* col_max, float* col_sum, float* diag) {
  if (img == nullptr || txt == nullptr || log_temp == nullptr || u == nullptr || v == nullptr ||
      inv_norm_img == nullptr || inv_norm_txt == nullptr || logit_scale_out == nullptr || S == nullptr ||
      row_lse == nullptr || col_max == nullptr || col_sum == nullptr || diag == nullptr || Bl <= 0 || Bg <= 0 ||
      E <= 0 || row_offset < 0 || row_offset + Bl > Bg)
    return FORDE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  l2norm_rows_kernel<<<(Bl + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(img),
                                                  static_cast<__nv_bfloat16*>(u), inv_norm_img, Bl, E, log_temp,
                                                  logit_scale_out);
  l2norm_rows_kernel<<<(Bg + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(txt),
                                                  static_cast<__nv_bfloat16*>(v), inv_norm_txt, Bg, E, nullptr,
                                                  nullptr);
  int rc = check_launch("l2norm_rows_kernel", 2);
  if (rc != FORDE_OK) return rc;
  // S[Bl][Bg] = exp(log_temp) * u . v^T  : both operands K-major (k = E)
  GemmArgs g{u, false, E, v, false, E, Bl, Bg, E, true, 0x1000000000000000ull, 0x1000000000000000ull};
  rc = gemm_f32(st, g, S, Bg, nullptr, nullptr, 1.f, logit_scale_out, false);
  if (rc != FORDE_OK) return rc;
  row_lse_kernel<<<(Bl + 7) / 8, 256, 0, st>>>(S, Bl, Bg, row_offset, row_lse, diag);
  col_lse_kernel<<<(Bg + 31) / 32, dim3(32, 8), 0, st>>>(S, Bl, Bg, col_max, col_sum);
  return check_launch("clip lse kernels", 2);
}

int forde_clip_col_rescale(void* stream, const float* col_max_local, const float* col_max_global, float* col_sum,
                           int Bg) {
  if (col_max_local == nullptr || col_max_global == nullptr || col_sum == nullptr || Bg <= 0) return FORDE_EINVAL;
  clip_col_rescale_kernel<<<(Bg + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      col_max_local, col_max_global, col_sum, Bg);
  return check_launch("clip_col_rescale_kernel");
}

int forde_clip_loss(void* stream, const float* row_lse, const float* diag, const float* col_max,
                    const float* col_sum, int Bl, int Bg, int row_offset, float* col_lse, float* loss) {
  if (row_lse == nullptr || diag == nullptr || col_max == nullptr || col_sum == nullptr || col_lse == nullptr ||
      loss == nullptr || Bl <= 0 || Bg <= 0)
    return FORDE_EINVAL;
  clip_loss_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_lse, diag, col_max, col_sum, Bl, Bg,
                                                                     row_offset, col_lse, loss);
  return check_launch("clip_loss_kernel");
}

int forde_clip_bwd(void* stream, const void* u, const void* v, const float* S, const float* row_lse,
                   const float* col_lse, const float* logit_scale, int Bl, int Bg, int E, int row_offset, void* dS,
                   float* dU, float* dV, float* d_log_temp) {
  if (u == nullptr || v == nullptr || S == nullptr || row_lse == nullptr || col_lse == nullptr ||
      logit_scale == nullptr || dS == nullptr || dU == nullptr || dV == nullptr || Bl <= 0 || Bg <= 0 || E <= 0)
    return FORDE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const size_t total = static_cast<size_t>(Bl) * Bg;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 8 * sms) blocks = 8 * sms;
  const int ld_ds = (Bg + 7) / 8 * 8;  // dS row pitch: TMA needs 16-byte pitches
  clip_ds_kernel<<<blocks, 256, 0, st>>>(S, row_lse, col_lse, Bl, Bg, row_offset, static_cast<__nv_bfloat16*>(dS),
                                         ld_ds, d_log_temp);
  int rc = check_launch("clip_ds_kernel");
  if (rc != FORDE_OK) return rc;
  // dU[Bl][E] = scale * dS . v   : A = dS [Bl][Bg] K-major ; B = v [Bg][E] as (n = e, k = j) MN-major
  GemmArgs gu{dS, false, ld_ds, v, true, E, Bl, E, Bg, true, 0x1000000000000000ull, 0x1000000000000000ull};
  rc = gemm_f32(st, gu, dU, E, nullptr, nullptr, 1.f, logit_scale, false);
  if (rc != FORDE_OK) return rc;
  // dV[Bg][E] = scale * dS^T . u : A = dS as (m = j, k = i) MN-major ; B = u [Bl][E] as (n = e, k = i) MN-major
  GemmArgs gv{dS, true, ld_ds, u, true, E, Bg, E, Bl, true, 0x1000000000000000ull, 0x1000000000000000ull};
  return gemm_f32(st, gv, dV, E, nullptr, nullptr, 1.f, logit_scale, false);
}

int forde_l2norm_bwd(void* stream, const void* u, const float* inv_norm, const float* dU, float* d_x, int B,
                     int E) {
  if (u == nullptr || inv_norm == nullptr || dU == nullptr || d_x == nullptr || B <= 0 || E <= 0)
    return FORDE_EINVAL;
  l2norm_bwd_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(u), inv_norm, dU, d_x, B, E);
  return check_launch("l2norm_bwd_kernel");
}

}  // extern "C"

// ------------------------------------------------------------------ top-k activation sparsification
// Optional (off by default): per token keep the k largest |h| over the H neurons, zero the rest.  hT is neuron-major
// [H][ld]; a block stages 16 token columns x H neurons in shared memory, one warp per token then
//   1. finds the k-th largest magnitude by a 15-step bitwise radix search on the bf16 magnitude bits,
//   2. keeps everything above it and the lowest-indexed ties (jax.lax.top_k tie rule), zeroes the rest IN THE STAGED
//      TILE and records a kept-bitmask [ceil(H/32)][M] (consumed by the dz epilogue of the backward GEMM: no
//      masking pass over dz^T, and the gradient statistics are taken of the masked dL/dz directly),
//   3. writes the kept indices ordered like lax.top_k (|value| descending, equal -> lower index first) and the
//      per-token Hoyer sparsity of the dense h (sensing.py:5-32 formula over the H axis);
// the block then writes the masked tile back with the access pattern it was read with (32-byte runs per neuron row):
// one read and one write of h^T in total, no scattered 2-byte read-modify-writes.
namespace forde {

constexpr int TK_TOK = 16;  // tokens per block (one warp each)

__global__ void __launch_bounds__(TK_TOK * 32)
topk_sparsify_kernel(__nv_bfloat16* __restrict__ hT, long long ld, int M, int H, int k, int* __restrict__ idx_out,
                     float* __restrict__ hoyer_out, uint32_t* __restrict__ keep_bits) {
  extern __shared__ uint16_t tile[];  // [TK_TOK][H] raw bf16 bits, then [TK_TOK][k] int scratch
  int* scratch_all = reinterpret_cast<int*>(tile + static_cast<size_t>(TK_TOK) * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * TK_TOK;
  uint16_t* h16 = reinterpret_cast<uint16_t*>(hT);
  for (int i = threadIdx.x; i < H * TK_TOK; i += blockDim.x) {  // TK_TOK consecutive tokens of every neuron row
    const int h = i / TK_TOK, t = i - h * TK_TOK;
    const int m = m0 + t;
    tile[t * H + h] = (m < M) ? h16[static_cast<size_t>(h) * ld + m] : 0;
  }
  __syncthreads();
  const int m = m0 + warp;
  if (m < M) {
    uint16_t* col = tile + warp * H;
    int* scratch = scratch_all + warp * k;
    float l1 = 0.f, l2 = 0.f;
    for (int h = lane; h < H; h += 32) {
      const float v = __uint_as_float(static_cast<uint32_t>(col[h]) << 16);
      l1 += fabsf(v);
      l2 = fmaf(v, v, l2);
    }
    l1 = warp_sum(l1);
    l2 = sqrtf(warp_sum(l2));
    if (lane == 0 && hoyer_out != nullptr) {
      const float sn = sqrtf(static_cast<float>(H));
      hoyer_out[m] = (H == 1 || l2 == 0.f) ? 0.f : (sn - l1 / l2) / (sn - 1.f);
    }
    // 1. radix search: after the loop `thr` is the key of the k-th largest magnitude and `need` of the elements
    //    equal to it are kept
    uint32_t thr = 0;
    int need = k;
    for (int bit = 14; bit >= 0; --bit) {
      const uint32_t cand = thr | (1u << bit);
      const uint32_t hi_mask = (~((1u << bit) - 1u)) & 0x7FFFu;
      int cnt = 0;
      for (int h = lane; h < H; h += 32) cnt += (((static_cast<uint32_t>(col[h]) & 0x7FFFu) & hi_mask) == cand) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (cnt >= need) thr = cand;
      else need -= cnt;
    }
    // 2. keep mask and the unsorted list of kept indices (ascending index) in shared scratch.  The ranking of step 3
    //    needs the kept magnitudes, so dropped entries are zeroed in the tile only after it (`drop` remembers them:
    //    one bit per 32-neuron group and lane).
    int n_kept = 0, n_tie = 0;
    for (int h0 = 0; h0 < H; h0 += 32) {
      const int h = h0 + lane;
      const uint32_t key = (h < H) ? (static_cast<uint32_t>(col[h]) & 0x7FFFu) : 0u;
      const bool above = (h < H) && key > thr;
      const bool tie = (h < H) && key == thr;
      const unsigned bt = __ballot_sync(0xffffffffu, tie);
      const bool keep = above || (tie && (n_tie + __popc(bt & ((1u << lane) - 1u))) < need);
      const unsigned bk = __ballot_sync(0xffffffffu, keep);
      if (keep) scratch[n_kept + __popc(bk & ((1u << lane) - 1u))] = h;
      if (keep_bits != nullptr && lane == 0) keep_bits[static_cast<size_t>(h0 >> 5) * M + m] = bk;
      n_kept += __popc(bk);
      n_tie += __popc(bt);
    }
    __syncwarp();
    // 3. lax.top_k order by counting rank: (|value| descending, index ascending); ranks are a permutation of 0..k-1
    if (idx_out != nullptr) {
      for (int a = lane; a < k; a += 32) {
        const int ia = scratch[a];
        const uint32_t ka = static_cast<uint32_t>(col[ia]) & 0x7FFFu;
        int rank = 0;
        for (int b = 0; b < k; ++b) {
          const int ib = scratch[b];
          const uint32_t kb = static_cast<uint32_t>(col[ib]) & 0x7FFFu;
          rank += (kb > ka || (kb == ka && ib < ia)) ? 1 : 0;
        }
        idx_out[static_cast<size_t>(m) * k + rank] = ia;
      }
    }
    __syncwarp();
    // 4. zero the dropped entries in the staged tile: everything below the threshold, and the ties beyond the first
    //    `need` (the same walk as step 2)
    n_tie = 0;
    for (int h0 = 0; h0 < H; h0 += 32) {
      const int h = h0 + lane;
      const uint32_t key = (h < H) ? (static_cast<uint32_t>(col[h]) & 0x7FFFu) : 0u;
      const bool tie = (h < H) && key == thr;
      const unsigned bt = __ballot_sync(0xffffffffu, tie);
      const bool keep = (h < H) && (key > thr || (tie && (n_tie + __popc(bt & ((1u << lane) - 1u))) < need));
      if (h < H && !keep) col[h] = 0;
      n_tie += __popc(bt);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * TK_TOK; i += blockDim.x) {  // masked tile back, same runs as the load
    const int h = i / TK_TOK, t = i - h * TK_TOK;
    const int mm = m0 + t;
    if (mm < M) h16[static_cast<size_t>(h) * ld + mm] = tile[t * H + h];
  }
}

}  // namespace forde

extern "C" {

int forde_sl_topk_sparsify(void* stream, void* hT, long long ldz, int M, int H, int k, int* idx_out, float* hoyer_out,
                           uint32_t* keep_bits) {
  if (hT == nullptr || M <= 0 || H <= 0 || k <= 0 || k > H || ldz < M) return FORDE_EINVAL;
  const int smem = forde::TK_TOK * H * 2 + forde::TK_TOK * k * 4;
  if (smem > 200 * 1024) return FORDE_EINVAL;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(forde::topk_sparsify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  forde::topk_sparsify_kernel<<<(M + forde::TK_TOK - 1) / forde::TK_TOK, forde::TK_TOK * 32, smem,
                                static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(hT), ldz, M, H, k,
                                                                     idx_out, hoyer_out, keep_bits);
  return forde::check_launch("topk_sparsify_kernel");
}

}  // extern "C"
