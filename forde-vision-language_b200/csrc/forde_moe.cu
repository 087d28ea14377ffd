// MoE variant of the FFN slot (BASELINE config 4; reference src/forde/moe.py:57-281): router + top-k + histogram +
// auxiliary loss + expert-usage statistics, device-side routing plan, and the expert MLPs as GROUPED tcgen05 GEMMs
// over the routed tokens only.  The reference runs every expert on every token and gathers the top-k
// (moe.py:163,169-185); evaluating only the selected experts gives the same result with k/E of the FLOPs.
// Nothing here synchronises with the host: counts, segment offsets and block->expert maps stay on the device.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "../../include/forde_b200.h"
#include "forde_host.h"

namespace forde {

constexpr int MOE_MAX_E = 16;
constexpr int MOE_MAX_K = 4;
constexpr int SEG_ALIGN = 128;  // expert segments are padded to the GEMM tile height

__device__ __forceinline__ float wsum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ router (moe.py:88-93,116-135,212-222,277-278)
// one warp per token: logits = x . W_r + b_r (fp32), probs = softmax(logits), top-k by value (descending, equal
// values -> lower index first: jax.lax.top_k), weights = softmax over the k selected logits.
// counts[E] (int, exact) via atomics; per-block sums of probs -> prob_partials[grid][E] (fixed order later).
template <int E>
__global__ void __launch_bounds__(256)
moe_router_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_r,
                  const float* __restrict__ b_r, int M, int d, int k, float* __restrict__ logits,
                  float* __restrict__ probs, int* __restrict__ topk_idx, float* __restrict__ topk_w,
                  int* __restrict__ counts, float* __restrict__ prob_partials) {
  extern __shared__ float w_s[];  // [d][E] fp32
  __shared__ float psum_s[8][E];
  __shared__ int cnt_s[E];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d * E; i += blockDim.x) w_s[i] = __bfloat162float(w_r[i]);
  if (threadIdx.x < E) cnt_s[threadIdx.x] = 0;
  __syncthreads();
  float psum[E];
#pragma unroll
  for (int e = 0; e < E; ++e) psum[e] = 0.f;
  for (int tok = blockIdx.x * 8 + warp; tok < M; tok += gridDim.x * 8) {
    const __nv_bfloat16* xr = x + static_cast<size_t>(tok) * d;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float xv = __bfloat162float(xr[i]);
      const float* wr = w_s + i * E;
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(xv, wr[e], acc[e]);
    }
    float lg[E];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      lg[e] = wsum32(acc[e]) + b_r[e];
      mx = fmaxf(mx, lg[e]);
    }
    float den = 0.f, pr[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      pr[e] = expf(lg[e] - mx);
      den += pr[e];
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      pr[e] *= inv;
      psum[e] += pr[e];
    }
    // top-k: repeated arg-max, strict '>' scanning ascending indices keeps the lowest index among equals
    int sel[MOE_MAX_K];
    float sv[MOE_MAX_K];
    unsigned taken = 0u;
#pragma unroll
    for (int j = 0; j < MOE_MAX_K; ++j) {
      if (j < k) {
        int best = -1;
        float bv = -INFINITY;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const bool free_e = ((taken >> e) & 1u) == 0u;
          if (free_e && (best < 0 || lg[e] > bv)) {
            best = e;
            bv = lg[e];
          }
        }
        sel[j] = best;
        sv[j] = bv;
        taken |= 1u << best;
      }
    }
    float wden = 0.f, wv[MOE_MAX_K];
#pragma unroll
    for (int j = 0; j < MOE_MAX_K; ++j) {
      if (j < k) {
        wv[j] = expf(sv[j] - sv[0]);
        wden += wv[j];
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        logits[static_cast<size_t>(tok) * E + e] = lg[e];
        probs[static_cast<size_t>(tok) * E + e] = pr[e];
      }
#pragma unroll
      for (int j = 0; j < MOE_MAX_K; ++j) {
        if (j < k) {
          topk_idx[static_cast<size_t>(tok) * k + j] = sel[j];
          topk_w[static_cast<size_t>(tok) * k + j] = wv[j] / wden;
          atomicAdd(&cnt_s[sel[j]], 1);
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < E; ++e) psum_s[warp][e] = psum[e];
  }
  __syncthreads();
  if (threadIdx.x < E) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += psum_s[w][threadIdx.x];
    prob_partials[static_cast<size_t>(blockIdx.x) * E + threadIdx.x] = t;
    atomicAdd(&counts[threadIdx.x], cnt_s[threadIdx.x]);
  }
}

// ------------------------------------------------------------------ routing plan (three small kernels)
// counts -> 128-aligned expert segments, block -> expert map, auxiliary loss (moe.py:212-222), expert-usage
// accumulation (moe.py:277-279) and a STABLE slot for every (token, j) entry (entries of an expert keep token order).
// The (token, j) entries are cut into chunks of PLAN_CHUNK; a chunk is one thread block:
//   1. moe_plan_count_kernel : per-chunk, per-expert entry counts (and the -1 fill of src_of_slot)
//   2. moe_plan_scan_kernel  : one block; segments, aux loss, usage, and the exclusive scan of the chunk counts
//                              -> chunk_base[chunk][e] = first slot of chunk's entries for expert e
//   3. moe_plan_assign_kernel: per chunk, rank of every entry among the chunk's entries of the same expert
//                              (warp match + prefix over the warps) + chunk_base -> slot
// (round 1 did all of it in one 1024-thread block, 16 serial passes at C4: 35 us of a 430 us slot.)
constexpr int PLAN_CHUNK = 1024;

__global__ void __launch_bounds__(PLAN_CHUNK)
moe_plan_count_kernel(const int* __restrict__ topk_idx, int total, int E, int r_cap, int* __restrict__ chunk_counts,
                      int* __restrict__ src_of_slot) {
  __shared__ int cnt_s[MOE_MAX_E];
  const int tid = threadIdx.x;
  if (tid < E) cnt_s[tid] = 0;
  __syncthreads();
  const int i = blockIdx.x * PLAN_CHUNK + tid;
  const int e = (i < total) ? topk_idx[i] : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, e);
  if (e >= 0 && (peers & ((1u << (tid & 31)) - 1u)) == 0u) atomicAdd(&cnt_s[e], __popc(peers));  // warp leader per expert
  __syncthreads();
  if (tid < E) chunk_counts[blockIdx.x * E + tid] = cnt_s[tid];
  for (int s = blockIdx.x * PLAN_CHUNK + tid; s < r_cap; s += gridDim.x * PLAN_CHUNK) src_of_slot[s] = -1;
}

__global__ void __launch_bounds__(256)
moe_plan_scan_kernel(const int* __restrict__ counts, const float* __restrict__ prob_partials, int n_partials, int M,
                     int k, int E, int r_cap, float aux_weight, int n_chunks, int* __restrict__ chunk_counts,
                     int* __restrict__ seg_offsets, int* __restrict__ n_blocks_dev, int* __restrict__ block_expert,
                     float* __restrict__ aux_loss, float* __restrict__ expert_usage, int* __restrict__ step_count) {
  __shared__ int seg_s[MOE_MAX_E + 1];
  __shared__ float aux_s[MOE_MAX_E];
  const int tid = threadIdx.x;
  if (tid == 0) {
    int off = 0;
    for (int e = 0; e < E; ++e) {
      seg_s[e] = off;
      off += (counts[e] + SEG_ALIGN - 1) / SEG_ALIGN * SEG_ALIGN;
    }
    seg_s[E] = off;
    *n_blocks_dev = off / SEG_ALIGN;
    if (step_count != nullptr) *step_count += 1;
  }
  if (tid < E) {
    double s = 0.0;
    for (int b = 0; b < n_partials; ++b) s += prob_partials[static_cast<size_t>(b) * E + tid];  // fixed order
    const float mean_p = static_cast<float>(s / static_cast<double>(M));
    if (expert_usage != nullptr) expert_usage[tid] += mean_p;
    aux_s[tid] = (static_cast<float>(counts[tid]) / static_cast<float>(M * k)) * mean_p;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f;
    for (int e = 0; e < E; ++e) a += aux_s[e];
    *aux_loss = a * static_cast<float>(E) * aux_weight;
  }
  if (tid <= E) seg_offsets[tid] = seg_s[tid];
  if (tid < E) {  // exclusive scan of this expert's chunk counts, in place: chunk_counts becomes chunk_base
    int run = seg_s[tid];
    for (int c = 0; c < n_chunks; ++c) {
      const int v = chunk_counts[c * E + tid];
      chunk_counts[c * E + tid] = run;
      run += v;
    }
  }
  for (int b = tid; b < r_cap / SEG_ALIGN; b += blockDim.x) {
    int e_of = -1;
    for (int e = 0; e < E; ++e)
      if (b * SEG_ALIGN >= seg_s[e] && b * SEG_ALIGN < seg_s[e + 1]) e_of = e;
    block_expert[b] = e_of;
  }
}

__global__ void __launch_bounds__(PLAN_CHUNK)
moe_plan_assign_kernel(const int* __restrict__ topk_idx, int total, int E, const int* __restrict__ chunk_base,
                       int* __restrict__ slot_of, int* __restrict__ src_of_slot) {
  __shared__ int wcnt_s[PLAN_CHUNK / 32][MOE_MAX_E];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * PLAN_CHUNK + tid;
  const int e = (i < total) ? topk_idx[i] : -1;
  for (int j = lane; j < E; j += 32) wcnt_s[warp][j] = 0;
  __syncwarp();
  const unsigned peers = __match_any_sync(0xffffffffu, e);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  if (e >= 0 && rank == 0) wcnt_s[warp][e] = __popc(peers);
  __syncthreads();
  if (tid < E) {  // exclusive prefix over the warps, per expert; afterwards wcnt_s holds each warp's first slot
    int run = chunk_base[blockIdx.x * E + tid];
    for (int w = 0; w < PLAN_CHUNK / 32; ++w) {
      const int c = wcnt_s[w][tid];
      wcnt_s[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
  if (e >= 0) {
    const int slot = wcnt_s[warp][e] + rank;
    slot_of[i] = slot;
    src_of_slot[slot] = i;
  }
}

// ------------------------------------------------------------------ row movers (HBM-bound, one warp per row)
// xg[slot] = x[src / k]  (zeros for padding slots)
__global__ void moe_gather_kernel(const uint4* __restrict__ x, const int* __restrict__ src_of_slot, int k,
                                  uint4* __restrict__ xg, int rows, int nvec) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int src = src_of_slot[row];
  for (int v = lane; v < nvec; v += 32)
    xg[static_cast<size_t>(row) * nvec + v] =
        (src >= 0) ? x[static_cast<size_t>(src / k) * nvec + v] : make_uint4(0u, 0u, 0u, 0u);
}

// dyg[slot] = w[src] * dout[src / k]  (zeros for padding slots)
__global__ void moe_scatter_kernel(const __nv_bfloat16* __restrict__ dout, const int* __restrict__ src_of_slot,
                                   const float* __restrict__ topk_w, int k, __nv_bfloat16* __restrict__ dyg,
                                   int rows, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int src = src_of_slot[row];
  const float w = (src >= 0) ? topk_w[src] : 0.f;
  const __nv_bfloat16* in = dout + static_cast<size_t>(src >= 0 ? src / k : 0) * d;
  for (int c = lane * 2; c < d; c += 64) {
    const __nv_bfloat162 v =
        (src >= 0) ? *reinterpret_cast<const __nv_bfloat162*>(in + c) : __floats2bfloat162_rn(0.f, 0.f);
    *reinterpret_cast<__nv_bfloat162*>(dyg + static_cast<size_t>(row) * d + c) =
        __floats2bfloat162_rn(w * __low2float(v), w * __high2float(v));
  }
}

// out[token] = sum_j w[token, j] * yg[slot_of[token, j]]   (w == nullptr: plain sum, used for dx)
__global__ void moe_combine_kernel(const __nv_bfloat16* __restrict__ yg, const int* __restrict__ slot_of,
                                   const float* __restrict__ topk_w, __nv_bfloat16* __restrict__ out, int M, int k,
                                   int d) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= M) return;
  for (int c = lane * 2; c < d; c += 64) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < k; ++j) {
      const int slot = slot_of[static_cast<size_t>(tok) * k + j];
      const float w = (topk_w != nullptr) ? topk_w[static_cast<size_t>(tok) * k + j] : 1.f;
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(yg + static_cast<size_t>(slot) * d + c);
      a0 = fmaf(w, __low2float(v), a0);
      a1 = fmaf(w, __high2float(v), a1);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + static_cast<size_t>(tok) * d + c) = __floats2bfloat162_rn(a0, a1);
  }
}

// d_w[token, j] = <dout[token], yg[slot_of[token, j]]>
__global__ void moe_gate_grad_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ yg,
                                     const int* __restrict__ slot_of, float* __restrict__ d_w, int entries, int k,
                                     int d) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= entries) return;
  const __nv_bfloat16* a = dout + static_cast<size_t>(i / k) * d;
  const __nv_bfloat16* b = yg + static_cast<size_t>(slot_of[i]) * d;
  float acc = 0.f;
  for (int c = lane * 2; c < d; c += 64) {
    const __nv_bfloat162 u = *reinterpret_cast<const __nv_bfloat162*>(a + c);
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(b + c);
    acc = fmaf(__low2float(u), __low2float(v), acc);
    acc = fmaf(__high2float(u), __high2float(v), acc);
  }
  acc = wsum32(acc);
  if (lane == 0) d_w[i] = acc;
}

// out[e][c] += sum of mat[row][c] over expert e's segment.  block = 32 x 8: a warp covers 256 columns with 16-byte
// loads, 8 row lanes; grid = (column blocks, experts, SEG_SPLIT row slices), slices combined with fp32 atomics.
constexpr int SEG_SPLIT = 8;
__global__ void moe_seg_colsum_kernel(const __nv_bfloat16* __restrict__ mat, const int* __restrict__ seg, int cols,
                                      float* __restrict__ out) {
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x, ty = threadIdx.y, e = blockIdx.y;
  const int c = blockIdx.x * 256 + tx * 8;
  const int s0 = seg[e], s1 = seg[e + 1];
  const int per = ((s1 - s0) + SEG_SPLIT - 1) / SEG_SPLIT;
  const int r0 = s0 + blockIdx.z * per, r1 = min(s1, r0 + per);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {  // cols is a multiple of 64 (checked by the caller), rows are 16-byte aligned
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(mat + static_cast<size_t>(r) * cols + c));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[2 * i] += __uint_as_float(w[i] << 16);
        s[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = s[i];
  __syncthreads();
  const int t = ty * 32 + tx;
  float acc = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) acc += red[y][t];
  const int col = blockIdx.x * 256 + t;
  if (col < cols && r1 > r0) atomicAdd(out + static_cast<size_t>(e) * cols + col, acc);
}

// ------------------------------------------------------------------ router backward
// One warp per token.  Gradient of the logits through (a) the top-k softmax weights (moe.py:133) given d_w =
// dL/d(topk_w), and (b) probs = softmax(logits) given g = dL/d(probs) -- the aux loss contributes
// d_aux * aux_weight * E / M * counts_e / (M k) per expert (moe.py:212-222; the histogram itself is not
// differentiable) plus an optional upstream d_probs.  Then dx[token] += dlogits . W_r^T (bf16, in place),
// db_r / dW_r are reduced per CTA in shared memory and added with fp32 atomics.
template <int E>
__global__ void __launch_bounds__(256)
moe_router_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_r,
                      const float* __restrict__ probs, const int* __restrict__ topk_idx,
                      const float* __restrict__ topk_w, const float* __restrict__ d_topk_w,
                      const int* __restrict__ counts, const float* __restrict__ d_aux,
                      const float* __restrict__ d_probs, float aux_weight, int M, int d, int k,
                      __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_r, float* __restrict__ db_r) {
  extern __shared__ float sm[];  // [d][E] W_r fp32, then one [d][E] dW_r accumulator PER WARP (lane-owned rows)
  float* w_s = sm;
  float* dw_s = sm + d * E * (1 + (threadIdx.x >> 5));
  __shared__ float db_s[E];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d * E; i += blockDim.x) w_s[i] = __bfloat162float(w_r[i]);
  for (int i = threadIdx.x; i < 8 * d * E; i += blockDim.x) sm[d * E + i] = 0.f;
  if (threadIdx.x < E) db_s[threadIdx.x] = 0.f;
  __syncthreads();
  const float aux_scale = (d_aux != nullptr ? *d_aux : 0.f) * aux_weight * static_cast<float>(E) /
                          static_cast<float>(M) / static_cast<float>(M * k);
  for (int tok = blockIdx.x * 8 + warp; tok < M; tok += gridDim.x * 8) {
    float dl[E], pr[E], g[E];
    float pg = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      pr[e] = probs[static_cast<size_t>(tok) * E + e];
      g[e] = aux_scale * static_cast<float>(counts[e]) +
             (d_probs != nullptr ? d_probs[static_cast<size_t>(tok) * E + e] : 0.f);
      pg = fmaf(pr[e], g[e], pg);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) dl[e] = pr[e] * (g[e] - pg);
    float wd = 0.f;
    for (int j = 0; j < k; ++j)
      wd = fmaf(topk_w[static_cast<size_t>(tok) * k + j], d_topk_w[static_cast<size_t>(tok) * k + j], wd);
    for (int j = 0; j < k; ++j) {
      const int sel = topk_idx[static_cast<size_t>(tok) * k + j];
      const float v = topk_w[static_cast<size_t>(tok) * k + j] * (d_topk_w[static_cast<size_t>(tok) * k + j] - wd);
#pragma unroll
      for (int e = 0; e < E; ++e) dl[e] += (e == sel) ? v : 0.f;
    }
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) atomicAdd(&db_s[e], dl[e]);
    }
    __nv_bfloat16* dxr = dx + static_cast<size_t>(tok) * d;
    const __nv_bfloat16* xr = x + static_cast<size_t>(tok) * d;
    for (int i = lane; i < d; i += 32) {
      const float* wr = w_s + i * E;
      float acc = __bfloat162float(dxr[i]);
      const float xv = __bfloat162float(xr[i]);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        acc = fmaf(dl[e], wr[e], acc);
        dw_s[i * E + e] = fmaf(xv, dl[e], dw_s[i * E + e]);  // row i of this warp's accumulator belongs to this lane
      }
      dxr[i] = __float2bfloat16(acc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d * E; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[d * E * (1 + w) + i];
    atomicAdd(&dw_r[i], t);
  }
  if (threadIdx.x < E) atomicAdd(&db_r[threadIdx.x], db_s[threadIdx.x]);
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_moe_capacity_rows(int M, int k, int E) {
  if (M <= 0 || k <= 0 || E <= 0) return 0;
  // every expert segment is padded to 128 rows: at most E * 127 padding rows
  return (M * k + E * (SEG_ALIGN - 1) + SEG_ALIGN - 1) / SEG_ALIGN * SEG_ALIGN;
}

int forde_moe_num_prob_partials(int M) {
  if (M <= 0) return 0;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int need = (M + 7) / 8;
  return need < 4 * sms ? need : 4 * sms;
}

int forde_moe_plan_scratch_ints(int M, int k, int E) {
  if (M <= 0 || k <= 0 || E <= 0) return 0;
  return ((M * k + PLAN_CHUNK - 1) / PLAN_CHUNK) * E;
}

int forde_moe_route(void* stream, const void* x, const void* w_r, const float* b_r, int M, int d, int E, int k,
                    float aux_weight, float* logits, float* probs, int* topk_idx, float* topk_w, int* counts,
                    float* prob_partials, int* plan_scratch, int* seg_offsets, int* n_blocks_dev, int* block_expert,
                    int* slot_of, int* src_of_slot, float* aux_loss, float* expert_usage, int* step_count) {
  if (x == nullptr || w_r == nullptr || b_r == nullptr || logits == nullptr || probs == nullptr ||
      topk_idx == nullptr || topk_w == nullptr || counts == nullptr || prob_partials == nullptr ||
      plan_scratch == nullptr ||
      seg_offsets == nullptr || n_blocks_dev == nullptr || block_expert == nullptr || slot_of == nullptr ||
      src_of_slot == nullptr || aux_loss == nullptr || M <= 0 || d <= 0)
    return FORDE_EINVAL;
  if ((E != 4 && E != 8 && E != 16) || k < 1 || k > MOE_MAX_K || k > E) return FORDE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = forde_moe_num_prob_partials(M);
  const int smem = d * E * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) return FORDE_EINVAL;
  cudaMemsetAsync(counts, 0, sizeof(int) * E, st);
#define FORDE_ROUTER(EE)                                                                                         \
  do {                                                                                                           \
    static std::once_flag once;                                                                                  \
    std::call_once(once, [] {                                                                                    \
      cudaFuncSetAttribute(moe_router_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);      \
    });                                                                                                          \
    moe_router_kernel<EE><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(x),                         \
                                                   static_cast<const __nv_bfloat16*>(w_r), b_r, M, d, k, logits, \
                                                   probs, topk_idx, topk_w, counts, prob_partials);              \
  } while (0)
  if (E == 4) FORDE_ROUTER(4);
  else if (E == 8) FORDE_ROUTER(8);
  else FORDE_ROUTER(16);
#undef FORDE_ROUTER
  const int r_cap = forde_moe_capacity_rows(M, k, E);
  const int total = M * k;
  const int n_chunks = (total + PLAN_CHUNK - 1) / PLAN_CHUNK;
  moe_plan_count_kernel<<<n_chunks, PLAN_CHUNK, 0, st>>>(topk_idx, total, E, r_cap, plan_scratch, src_of_slot);
  moe_plan_scan_kernel<<<1, 256, 0, st>>>(counts, prob_partials, grid, M, k, E, r_cap, aux_weight, n_chunks,
                                          plan_scratch, seg_offsets, n_blocks_dev, block_expert, aux_loss,
                                          expert_usage, step_count);
  moe_plan_assign_kernel<<<n_chunks, PLAN_CHUNK, 0, st>>>(topk_idx, total, E, plan_scratch, slot_of, src_of_slot);
  return check_launch("moe_router_kernel + moe_plan kernels", 4);
}

int forde_moe_gather(void* stream, const void* x, const int* src_of_slot, int k, void* xg, int rows, int d) {
  if (x == nullptr || src_of_slot == nullptr || xg == nullptr || rows <= 0 || d <= 0 || k <= 0) return FORDE_EINVAL;
  if ((d & 7) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(xg) & 15) != 0)
    return FORDE_EALIGN;
  moe_gather_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), src_of_slot, k, static_cast<uint4*>(xg), rows, d / 8);
  return check_launch("moe_gather_kernel");
}

int forde_moe_scatter(void* stream, const void* dout, const int* src_of_slot, const float* topk_w, int k, void* dyg,
                      int rows, int d) {
  if (dout == nullptr || src_of_slot == nullptr || topk_w == nullptr || dyg == nullptr || rows <= 0 || d <= 0 ||
      k <= 0)
    return FORDE_EINVAL;
  if ((d & 1) != 0) return FORDE_EALIGN;
  moe_scatter_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), src_of_slot, topk_w, k, static_cast<__nv_bfloat16*>(dyg), rows, d);
  return check_launch("moe_scatter_kernel");
}

int forde_moe_combine(void* stream, const void* yg, const int* slot_of, const float* topk_w, void* out, int M, int k,
                      int d) {
  if (yg == nullptr || slot_of == nullptr || out == nullptr || M <= 0 || k <= 0 || d <= 0) return FORDE_EINVAL;
  if ((d & 1) != 0) return FORDE_EALIGN;
  moe_combine_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(yg), slot_of, topk_w, static_cast<__nv_bfloat16*>(out), M, k, d);
  return check_launch("moe_combine_kernel");
}

int forde_moe_gate_grad(void* stream, const void* dout, const void* yg, const int* slot_of, float* d_topk_w, int M,
                        int k, int d) {
  if (dout == nullptr || yg == nullptr || slot_of == nullptr || d_topk_w == nullptr || M <= 0 || k <= 0 || d <= 0)
    return FORDE_EINVAL;
  if ((d & 1) != 0) return FORDE_EALIGN;
  const int entries = M * k;
  moe_gate_grad_kernel<<<(entries + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(yg), slot_of, d_topk_w, entries, k,
      d);
  return check_launch("moe_gate_grad_kernel");
}

int forde_moe_router_bwd(void* stream, const void* x, const void* w_r, const float* probs, const int* topk_idx,
                         const float* topk_w, const float* d_topk_w, const int* counts, const float* d_aux,
                         const float* d_probs, float aux_weight, int M, int d, int E, int k, void* dx, float* dw_r,
                         float* db_r) {
  if (x == nullptr || w_r == nullptr || probs == nullptr || topk_idx == nullptr || topk_w == nullptr ||
      d_topk_w == nullptr || counts == nullptr || dx == nullptr || dw_r == nullptr || db_r == nullptr || M <= 0 ||
      d <= 0)
    return FORDE_EINVAL;
  if ((E != 4 && E != 8 && E != 16) || k < 1 || k > MOE_MAX_K) return FORDE_EINVAL;
  const int smem = 9 * d * E * static_cast<int>(sizeof(float));  // W_r + one dW_r accumulator per warp
  if (smem > 200 * 1024) return FORDE_EINVAL;                    // d * E <= 5688 (C4: 512 x 8 = 4096)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const int need = (M + 7) / 8;
  const int grid = need < sms ? need : sms;
#define FORDE_ROUTER_BWD(EE)                                                                                     \
  do {                                                                                                           \
    static std::once_flag once;                                                                                  \
    std::call_once(once, [] {                                                                                    \
      cudaFuncSetAttribute(moe_router_bwd_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  \
    });                                                                                                          \
    moe_router_bwd_kernel<EE><<<grid, 256, smem, st>>>(                                                          \
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_r), probs, topk_idx, topk_w,   \
        d_topk_w, counts, d_aux, d_probs, aux_weight, M, d, k, static_cast<__nv_bfloat16*>(dx), dw_r, db_r);     \
  } while (0)
  if (E == 4) FORDE_ROUTER_BWD(4);
  else if (E == 8) FORDE_ROUTER_BWD(8);
  else FORDE_ROUTER_BWD(16);
#undef FORDE_ROUTER_BWD
  return check_launch("moe_router_bwd_kernel");
}

// experts forward on the gathered rows: zg = xg . W_up[e] + b_up[e]; hg = gelu(zg); yg = hg . W_down[e] + b_down[e]
int forde_moe_experts_fwd(void* stream, const void* xg, const void* w_up_all, const float* b_up_all,
                          const void* w_down_all, const float* b_down_all, const int* n_blocks_dev,
                          const int* block_expert, void* zg, void* hg, void* yg, int r_cap, int d, int H, int E) {
  if (xg == nullptr || w_up_all == nullptr || w_down_all == nullptr || zg == nullptr || hg == nullptr ||
      yg == nullptr || r_cap <= 0 || (r_cap % SEG_ALIGN) != 0 || d <= 0 || H <= 0 || E <= 0)
    return FORDE_EINVAL;
  if ((d % 64) != 0 || (H % 64) != 0) return FORDE_EALIGN;  // expert slices must not share a k-block
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // up: A = xg [R][d] K-major ; B = W_up stacked [E*d][H] (k = d rows, n = H contiguous) MN-major, k offset e*d
  GemmArgs gu{xg, false, d, w_up_all, true, H, r_cap, H, d, false, 0x1000000000000000ull, 0x1000000000000000ull};
  GroupedM mu{n_blocks_dev, block_expert, d, 0, H, E};
  int rc = gemm_grouped_m(st, gu, mu, 1, zg, H, hg, nullptr, b_up_all);
  if (rc != FORDE_OK) return rc;
  // down: A = hg [R][H] K-major ; B = W_down stacked [E*H][d] MN-major, k offset e*H
  GemmArgs gd{hg, false, H, w_down_all, true, d, r_cap, d, H, false, 0x1000000000000000ull, 0x1000000000000000ull};
  GroupedM md{n_blocks_dev, block_expert, H, 0, d, E};
  return gemm_grouped_m(st, gd, md, 0, yg, d, nullptr, nullptr, b_down_all);
}

// experts backward: dzg = (dyg . W_down[e]^T) * gelu'(zg); dxg = dzg . W_up[e]^T;
// dW_up[e] = xg_e^T dzg_e ; dW_down[e] = hg_e^T dyg_e ; db_up[e] += colsum(dzg_e) ; db_down[e] += colsum(dyg_e)
int forde_moe_experts_bwd(void* stream, const void* xg, const void* dyg, const void* w_up_all,
                          const void* w_down_all, const void* zg, const void* hg, const int* n_blocks_dev,
                          const int* block_expert, const int* seg_offsets, void* dzg, void* dxg, float* dw_up_all,
                          float* db_up_all, float* dw_down_all, float* db_down_all, int r_cap, int d, int H, int E) {
  if (xg == nullptr || dyg == nullptr || w_up_all == nullptr || w_down_all == nullptr || zg == nullptr ||
      hg == nullptr || dzg == nullptr || dxg == nullptr || dw_up_all == nullptr || dw_down_all == nullptr ||
      db_up_all == nullptr || db_down_all == nullptr || r_cap <= 0 || (r_cap % SEG_ALIGN) != 0)
    return FORDE_EINVAL;
  if ((d % 64) != 0 || (H % 64) != 0) return FORDE_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // dz: A = dyg [R][d] K-major ; B = W_down stacked [E*H][d] as rows n = neuron (K-major), n offset e*H
  GemmArgs gz{dyg, false, d, w_down_all, false, d, r_cap, H, d, false, 0x1000000000000000ull, 0x1000000000000000ull};
  GroupedM mz{n_blocks_dev, block_expert, 0, H, 0, E};
  int rc = gemm_grouped_m(st, gz, mz, 2, dzg, H, nullptr, zg, nullptr);
  if (rc != FORDE_OK) return rc;
  // dx: A = dzg [R][H] K-major ; B = W_up stacked [E*d][H] as rows n = input feature (K-major), n offset e*d
  GemmArgs gx{dzg, false, H, w_up_all, false, H, r_cap, d, H, false, 0x1000000000000000ull, 0x1000000000000000ull};
  GroupedM mx{n_blocks_dev, block_expert, 0, d, 0, E};
  rc = gemm_grouped_m(st, gx, mx, 0, dxg, d, nullptr, nullptr, nullptr);
  if (rc != FORDE_OK) return rc;
  // dW_up[e] (d x H) : A = xg [rows][d] MN-major (m = d) ; B = dzg [rows][H] MN-major (n = H) ; k = rows of segment e
  GemmArgs gwu{xg, true, d, dzg, true, H, d, H, r_cap, false, 0x1000000000000000ull, 0x1000000000000000ull};
  rc = gemm_grouped_k(st, gwu, seg_offsets, E, dw_up_all, H, static_cast<long long>(d) * H);
  if (rc != FORDE_OK) return rc;
  // dW_down[e] (H x d) : A = hg [rows][H] MN-major (m = H) ; B = dyg [rows][d] MN-major (n = d)
  GemmArgs gwd{hg, true, H, dyg, true, d, H, d, r_cap, false, 0x1000000000000000ull, 0x1000000000000000ull};
  rc = gemm_grouped_k(st, gwd, seg_offsets, E, dw_down_all, d, static_cast<long long>(d) * H);
  if (rc != FORDE_OK) return rc;
  moe_seg_colsum_kernel<<<dim3((H + 255) / 256, E, SEG_SPLIT), dim3(32, 8), 0, st>>>(static_cast<const __nv_bfloat16*>(dzg),
                                                                       seg_offsets, H, db_up_all);
  moe_seg_colsum_kernel<<<dim3((d + 255) / 256, E, SEG_SPLIT), dim3(32, 8), 0, st>>>(static_cast<const __nv_bfloat16*>(dyg),
                                                                       seg_offsets, d, db_down_all);
  return check_launch("moe_seg_colsum_kernel", 2);
}

}  // extern "C"
