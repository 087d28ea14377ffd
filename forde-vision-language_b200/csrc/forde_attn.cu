// Self-attention core for SHORT sequences (T <= 128, head_dim 64): the text tower of the FORDE dual encoder
// (README.md:30-34,42-43: "self-attention -> add/norm -> StatefulLayer -> add/norm"; T = 77 tokens).
//
// Why it exists: for T = 77 every library kernel in the image (cuDNN's sm80 wmma fallback, FlashAttention-2, torch's
// flash / memory-efficient backends) needs 275-355 us per layer for forward + backward -- 11 GFLOP and 160 MB of
// traffic (profiles/r2_sdpa_probe.txt) -- because they tile the sequence for long contexts.  Here the WHOLE sequence of
// one (batch, head) is one tcgen05 tile: Q, K, V (and dO) are loaded once by TMA, scores live in TMEM, and the
// probabilities are staged once in shared memory in the 128-byte-swizzled layout that the tensor core reads both as a
// K-major operand (P . V, dS . K) and as an MN-major operand (P^T . dO, dS^T . Q) -- no transposes, no second pass.
//
//   forward : S = Q K^T (M = 128 query rows padded, N = Tpad keys, K = 64)           -> TMEM cols [0, 128)
//             one thread per query row: p = exp2(S * scale * log2e - max), row sum   -> P (bf16) in shared memory
//             O = P V (K = Tpad)                                                      -> TMEM cols [128, 192)
//             o = O / rowsum -> bf16 -> the (idle) P buffer -> one TMA store ; lse = ln sum_j exp(scale * s_j)
//   backward: S = Q K^T, dP = dO V^T                                                  -> TMEM [0,128), [128,256)
//             p = exp2(S * scale * log2e - lse * log2e) ; D = sum_j p dP ; dS = p (dP - D) scale   (two passes over
//             TMEM instead of keeping 256 values per thread)                         -> P, dS (bf16) in shared memory
//             dV = P^T dO, dK = dS^T Q, dQ = dS K                                     -> TMEM [0,64), [64,128), [128,192)
//             each gradient -> bf16 rows in a swizzled staging tile -> one TMA store (rows beyond T clipped)
//             (operand tiles double buffered: the next head's 64 KB load while this one computes)
// What bounds it: latency.  Five warps per CTA walk an item's phases one after the other (ncu: 8 % occupancy, MUFU
// 13 %, tensor 6 %, profiles/r2_ncu_full_aux_kernels.md); a backward item reads S and dP twice and the three gradients
// once from TMEM (~300 KB = 2.3 k clocks at the measured 128 B/clk/SM) of the ~10 k clocks it takes.  More CTAs per SM
// do not help (measured in round 2 with compact Tpad-row tiles, 2-3 CTAs per SM: forward 38 -> 65 us, backward 101 ->
// 108 us per text layer, reverted) and a second item in flight per CTA does not fit (operands + P + dS = 128 KB per
// item).  Double-buffering the operand loads (107 -> 101 us) and writing dQ / dK / dV through staged TMA stores
// instead of per-thread 128-byte rows (104 -> 93 us, same box) are kept; so are one MUFU.EX2 per exponential instead
// of exp2f's rescaled one and mask-free arithmetic for the chunks that lie entirely below T (93 -> 74 us, forward
// 45 -> 41 us: the kernels are issue / latency bound, so instructions count).  Shipped: forward 41 us + backward 74 us
// per text layer at C2 (B = 256, 8 heads) against 51 + 195 us for cuDNN's kernels.
// Rows / keys beyond T: TMA zero-fills them (3-D tensor map {columns, tokens, batch}), masked keys get p = 0, and the
// padded query rows of P / dS are written as zeros because they are reduction (K) indices of dV and dK.
// Roles: warps 0-3 = one thread per row (TMEM lane quarter = warp), warp 4 = TMA + MMA issue (one thread).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "../../include/forde_b200.h"
#include "forde_host.h"
#include "sm100_ptx.cuh"

namespace forde {

constexpr int AT_ROWS = 128;                     // tile rows (queries or keys), >= T
constexpr int AT_DH = 64;                        // head dimension (one 128-byte swizzled row)
constexpr int AT_TILE = AT_ROWS * AT_DH * 2;     // 16 KB
constexpr int AT_THREADS = 160;
constexpr int AT_TMEM_COLS = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV, tmDO;
  CUtensorMap tmO, tmDQ, tmDK, tmDV;   // outputs leave through shared memory and TMA stores (rows beyond T clipped)
  __nv_bfloat16* o;      // forward output [B*T][ld]
  float* lse;            // [B][heads][T]
  __nv_bfloat16 *dq, *dk, *dv;
  int B, T, heads, Tpad;
  long long ld;
  float scale;           // 1 / sqrt(head_dim)
};

// shared address of the 16-byte chunk holding columns [col, col + 8) of row r in a [2 tiles][128 rows][64 cols]
// bf16 buffer with TMA's 128-byte swizzle (tile = col / 64)
__device__ __forceinline__ uint32_t sw128_chunk(uint32_t base, int r, int col) {
  return base + static_cast<uint32_t>(col >> 6) * AT_TILE + static_cast<uint32_t>(r) * 128u +
         ((static_cast<uint32_t>((col & 63) >> 3) ^ static_cast<uint32_t>(r & 7)) << 4);
}

__device__ __forceinline__ float at_ex2(float x) {   // one MUFU.EX2 (exp2f adds a denormal-range rescale per element)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void at_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void at_tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 64 fp32 accumulator values of row r (two 32-column TMEM loads) -> 64 bf16 in a [128 rows][64 cols] staging tile with
// TMA's 128-byte swizzle.  (Every thread storing its own 128-byte row to global memory -- 32 lines per store
// instruction -- is what the first version of these kernels spent most of an item on.)
__device__ __forceinline__ void stage_row64(uint32_t tile_u, int r, const uint32_t* v0, const uint32_t* v1, float mul) {
  const uint32_t row = tile_u + static_cast<uint32_t>(r) * 128u, rx = static_cast<uint32_t>(r & 7);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t* v = half == 0 ? v0 : v1;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4(row + ((static_cast<uint32_t>(half * 4 + j) ^ rx) << 4),
                   pack_bf16x2(__uint_as_float(v[8 * j + 0]) * mul, __uint_as_float(v[8 * j + 1]) * mul),
                   pack_bf16x2(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul),
                   pack_bf16x2(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul),
                   pack_bf16x2(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul));
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 2 * AT_TILE;
  uint8_t* sP = smem + 3 * AT_TILE;                      // [2][128][64] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * AT_TILE);
  uint64_t* bar_load = bars;       // TMA bytes of Q, K, V
  uint64_t* bar_s = bars + 1;      // S = Q K^T complete
  uint64_t* bar_p = bars + 2;      // P written by the 128 row threads
  uint64_t* bar_o = bars + 3;      // O = P V complete
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 128) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmO);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr_s, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int items = p.B * p.heads;
  const int T = p.T, Tpad = p.Tpad;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(AT_ROWS, Tpad, false, false);   // Q (K-major) x K (K-major)
      const uint32_t idesc_o = make_idesc_bf16(AT_ROWS, AT_DH, false, true);   // P (K-major) x V (MN-major)
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1u) {
        const int b = it / p.heads, h = it - b * p.heads;
        mbar_expect_tx(bar_load, 3 * AT_TILE);
        tma_load_3d(&p.tmQ, bar_load, sQ, h * AT_DH, 0, b);
        tma_load_3d(&p.tmK, bar_load, sK, h * AT_DH, 0, b);
        tma_load_3d(&p.tmV, bar_load, sV, h * AT_DH, 0, b);
        mbar_wait(bar_load, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024),
                    make_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, ph);
        tc_fence_after();
        for (int ks = 0; ks < Tpad / 16; ++ks)   // reduction over the keys, 16 per step
          umma_bf16(tmem_base + 128,
                    make_smem_desc_sw128(smem_u32(sP) + (ks >> 2) * AT_TILE + (ks & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(smem_u32(sV) + ks * 2048, AT_TILE, 1024), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(bar_o);
        mbar_wait(bar_o, ph);   // P and V are free again: the next item's loads overlap this item's output stores
      }
    }
  } else {
    const int r = warp * 32 + lane;                                    // query row = TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t sP_u = smem_u32(sP);
    const float sl2 = p.scale * kLog2e;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1u) {
      const int b = it / p.heads, h = it - b * p.heads;
      mbar_wait(bar_s, ph);
      tc_fence_after();
      float mx = -INFINITY;
      for (int c0 = 0; c0 < Tpad; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_wait_ld();
        if (c0 + 32 <= T) {                      // chunk entirely below T: no per-element mask, four chains
          float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      mx *= sl2;                                 // sl2 > 0: max(s) * sl2 = max(s * sl2)
      // the P buffer also carried the previous item's O on its way out: that TMA store must have read it
      if (it != static_cast<int>(blockIdx.x)) {
        if (threadIdx.x == 0) tma_wait_read<0>();
        at_bar_sync(1, 128);
      }
      float sum = 0.f;
      for (int c0 = 0; c0 < Tpad; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_wait_ld();
        const bool full = c0 + 32 <= T;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = c0 + j8 * 8 + 2 * i;
            float p0 = at_ex2(fmaf(__uint_as_float(v[j8 * 8 + 2 * i]), sl2, -mx));
            float p1 = at_ex2(fmaf(__uint_as_float(v[j8 * 8 + 2 * i + 1]), sl2, -mx));
            if (!full) {                         // (warp-uniform) the chunk straddles T: keys beyond T get p = 0
              p0 = (c < T) ? p0 : 0.f;
              p1 = (c + 1 < T) ? p1 : 0.f;
            }
            sum += p0 + p1;
            w[i] = pack_bf16x2(p0, p1);
          }
          if (c0 + j8 * 8 < Tpad) st_shared_v4(sw128_chunk(sP_u, r, c0 + j8 * 8), w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
      if (r < T) p.lse[(static_cast<size_t>(b) * p.heads + h) * T + r] = (mx + log2f(sum)) * kLn2;
      mbar_wait(bar_o, ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(taddr + 128, o0);
      tmem_ld32(taddr + 160, o1);
      tmem_wait_ld();
      tc_fence_before();
      stage_row64(sP_u, r, o0, o1, 1.f / sum);   // P V has completed: tile 0 of the P buffer is idle
      fence_proxy_async_smem();
      at_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        at_tma_store_3d(&p.tmO, sP_u, h * AT_DH, 0, b);
        tma_commit();
      }
    }
    if (threadIdx.x == 0) tma_wait_all<0>();     // the last store has landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // operand tiles Q, K, V, dO are double buffered: the next (batch, head)'s 64 KB are in flight while this one computes
  uint8_t* sOp = smem;                                   // [2 stages][Q, K, V, dO] x 16 KB
  uint8_t* sP = smem + 8 * AT_TILE;                      // [2][128][64] bf16
  uint8_t* sDS = smem + 10 * AT_TILE;                    // [2][128][64] bf16
  uint8_t* sOut = smem + 12 * AT_TILE;                   // [2] gradient tiles on their way out (TMA stores)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 14 * AT_TILE);
  uint64_t* bar_load = bars;       // [2] TMA bytes of Q, K, V, dO of a stage
  uint64_t* bar_sdp = bars + 2;    // S and dP complete
  uint64_t* bar_pds = bars + 3;    // P and dS written by the 128 row threads
  uint64_t* bar_g = bars + 4;      // dV, dK, dQ complete
  uint64_t* bar_epi = bars + 5;    // the 128 row threads have read dV, dK, dQ out of TMEM
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 128) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmDO);
    tma_prefetch_desc(&p.tmDQ);
    tma_prefetch_desc(&p.tmDK);
    tma_prefetch_desc(&p.tmDV);
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 128);
    mbar_init(bar_g, 1);
    mbar_init(bar_epi, 128);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr_s, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int items = p.B * p.heads;
  const int T = p.T, Tpad = p.Tpad;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(AT_ROWS, Tpad, false, false);   // Q x K, dO x V (all K-major)
      const uint32_t idesc_tt = make_idesc_bf16(AT_ROWS, AT_DH, true, true);   // P^T x dO, dS^T x Q (both MN-major)
      const uint32_t idesc_kt = make_idesc_bf16(AT_ROWS, AT_DH, false, true);  // dS (K-major) x K (MN-major)
      auto issue_loads = [&](int item, int stage) {
        const int b = item / p.heads, h = item - b * p.heads;
        uint8_t* s0 = sOp + stage * 4 * AT_TILE;
        mbar_expect_tx(&bar_load[stage], 4 * AT_TILE);
        tma_load_3d(&p.tmQ, &bar_load[stage], s0, h * AT_DH, 0, b);
        tma_load_3d(&p.tmK, &bar_load[stage], s0 + AT_TILE, h * AT_DH, 0, b);
        tma_load_3d(&p.tmV, &bar_load[stage], s0 + 2 * AT_TILE, h * AT_DH, 0, b);
        tma_load_3d(&p.tmDO, &bar_load[stage], s0 + 3 * AT_TILE, h * AT_DH, 0, b);
      };
      uint32_t ph = 0;
      int n = 0;
      if (static_cast<int>(blockIdx.x) < items) issue_loads(blockIdx.x, 0);
      for (int it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1u, ++n) {
        const int stage = n & 1;
        // the other stage was last read by the MMAs of item n - 1, which completed (bar_g) before this point
        if (it + static_cast<int>(gridDim.x) < items) issue_loads(it + gridDim.x, stage ^ 1);
        uint8_t* sQ = sOp + stage * 4 * AT_TILE;
        uint8_t* sK = sQ + AT_TILE;
        uint8_t* sV = sQ + 2 * AT_TILE;
        uint8_t* sDO = sQ + 3 * AT_TILE;
        mbar_wait(&bar_load[stage], (n >> 1) & 1);
        if (n > 0) mbar_wait(bar_epi, ph ^ 1u);   // the previous item's gradients have left TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k)      // S = Q K^T
          umma_bf16(tmem_base, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 0, 1024),
                    make_smem_desc_sw128(smem_u32(sK) + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k)      // dP = dO V^T
          umma_bf16(tmem_base + 128, make_smem_desc_sw128(smem_u32(sDO) + k * 32, 0, 1024),
                    make_smem_desc_sw128(smem_u32(sV) + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_sdp);
        mbar_wait(bar_pds, ph);
        tc_fence_after();
        for (int ks = 0; ks < Tpad / 16; ++ks) {  // reductions over the query rows (dV, dK) and over the keys (dQ)
          const uint32_t acc = ks > 0 ? 1u : 0u;
          // dV[key][:] = sum_q P[q][key] dO[q][:]
          umma_bf16(tmem_base, make_smem_desc_sw128(smem_u32(sP) + ks * 2048, AT_TILE, 1024),
                    make_smem_desc_sw128(smem_u32(sDO) + ks * 2048, AT_TILE, 1024), idesc_tt, acc);
          // dK[key][:] = sum_q dS[q][key] Q[q][:]
          umma_bf16(tmem_base + 64, make_smem_desc_sw128(smem_u32(sDS) + ks * 2048, AT_TILE, 1024),
                    make_smem_desc_sw128(smem_u32(sQ) + ks * 2048, AT_TILE, 1024), idesc_tt, acc);
          // dQ[q][:] = sum_key dS[q][key] K[key][:]
          umma_bf16(tmem_base + 128,
                    make_smem_desc_sw128(smem_u32(sDS) + (ks >> 2) * AT_TILE + (ks & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(smem_u32(sK) + ks * 2048, AT_TILE, 1024), idesc_kt, acc);
        }
        umma_commit(bar_g);
        mbar_wait(bar_g, ph);   // this stage's operand tiles and P / dS are free again
      }
    }
  } else {
    const int r = warp * 32 + lane;                                    // query row (S, dP, dQ) / key row (dK, dV)
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t sP_u = smem_u32(sP), sDS_u = smem_u32(sDS), sOut_u = smem_u32(sOut);
    const float sl2 = p.scale * kLog2e;
    uint32_t ph = 0;
    int stage = 0;                                                     // staging tiles alternate
    for (int it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1u) {
      const int b = it / p.heads, h = it - b * p.heads;
      // padded query rows (r >= T) are reduction indices of dV / dK: lse = +inf makes their p (and dS) exact zeros
      const float lse2 = (r < T) ? p.lse[(static_cast<size_t>(b) * p.heads + h) * T + r] * kLog2e : INFINITY;
      mbar_wait(bar_sdp, ph);
      tc_fence_after();
      float D = 0.f;                                                   // sum_j p_j dP_j  (= rowsum(dO * O))
      for (int c0 = 0; c0 < Tpad; c0 += 32) {
        uint32_t s[32], d[32];
        tmem_ld32(taddr + c0, s);
        tmem_ld32(taddr + 128 + c0, d);
        tmem_wait_ld();
        if (c0 + 32 <= T) {                      // chunk entirely below T: no per-element mask, four chains
          float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            d0 = fmaf(at_ex2(fmaf(__uint_as_float(s[j]), sl2, -lse2)), __uint_as_float(d[j]), d0);
            d1 = fmaf(at_ex2(fmaf(__uint_as_float(s[j + 1]), sl2, -lse2)), __uint_as_float(d[j + 1]), d1);
            d2 = fmaf(at_ex2(fmaf(__uint_as_float(s[j + 2]), sl2, -lse2)), __uint_as_float(d[j + 2]), d2);
            d3 = fmaf(at_ex2(fmaf(__uint_as_float(s[j + 3]), sl2, -lse2)), __uint_as_float(d[j + 3]), d3);
          }
          D += (d0 + d1) + (d2 + d3);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < T) D = fmaf(at_ex2(fmaf(__uint_as_float(s[j]), sl2, -lse2)), __uint_as_float(d[j]), D);
        }
      }
      for (int c0 = 0; c0 < Tpad; c0 += 32) {
        uint32_t s[32], d[32];
        tmem_ld32(taddr + c0, s);
        tmem_ld32(taddr + 128 + c0, d);
        tmem_wait_ld();
        const bool full = c0 + 32 <= T;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t pw[4], dw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = c0 + j8 * 8 + 2 * i, j = j8 * 8 + 2 * i;
            float p0 = at_ex2(fmaf(__uint_as_float(s[j]), sl2, -lse2));
            float p1 = at_ex2(fmaf(__uint_as_float(s[j + 1]), sl2, -lse2));
            if (!full) {                         // (warp-uniform) the chunk straddles T: keys beyond T get p = 0
              p0 = (c < T) ? p0 : 0.f;
              p1 = (c + 1 < T) ? p1 : 0.f;
            }
            pw[i] = pack_bf16x2(p0, p1);
            dw[i] = pack_bf16x2(p0 * (__uint_as_float(d[j]) - D) * p.scale, p1 * (__uint_as_float(d[j + 1]) - D) * p.scale);
          }
          if (c0 + j8 * 8 < Tpad) {
            st_shared_v4(sw128_chunk(sP_u, r, c0 + j8 * 8), pw[0], pw[1], pw[2], pw[3]);
            st_shared_v4(sw128_chunk(sDS_u, r, c0 + j8 * 8), dw[0], dw[1], dw[2], dw[3]);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
      mbar_wait(bar_g, ph);
      tc_fence_after();
      // dQ, dK, dV: TMEM -> bf16 rows in a swizzled staging tile -> one TMA store each (two tiles, alternating; rows
      // beyond T are clipped by the tensor map)
      auto read_out = [&](uint32_t tcol, const CUtensorMap* tm, bool last) {
        if (threadIdx.x == 0) tma_wait_read<1>();      // the store that used this tile two read-outs ago has read it
        at_bar_sync(1, 128);
        uint32_t a0[32], a1[32];
        tmem_ld32(taddr + tcol, a0);
        tmem_ld32(taddr + tcol + 32, a1);
        tmem_wait_ld();
        if (last) {                                    // the gradients have left TMEM: the next item's S / dP may start
          tc_fence_before();
          mbar_arrive(bar_epi);
        }
        const uint32_t tile_u = sOut_u + stage * AT_TILE;
        stage_row64(tile_u, r, a0, a1, 1.f);
        fence_proxy_async_smem();
        at_bar_sync(1, 128);
        if (threadIdx.x == 0) {
          at_tma_store_3d(tm, tile_u, h * AT_DH, 0, b);
          tma_commit();
        }
        stage ^= 1;
      };
      read_out(128, &p.tmDQ, false);
      read_out(64, &p.tmDK, false);
      read_out(0, &p.tmDV, true);
    }
    if (threadIdx.x == 0) tma_wait_all<0>();           // the last stores have landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

constexpr int attn_fwd_smem() { return 5 * AT_TILE + 256 + 1024; }
constexpr int attn_bwd_smem() { return 14 * AT_TILE + 256 + 1024; }

static int fill_attn(AttnParams& p, const void* q, const void* k, const void* v, const void* d_o, int B, int T,
                     int heads, long long ld, void* o = nullptr, void* dq = nullptr, void* dk = nullptr,
                     void* dv = nullptr) {
  if (q == nullptr || k == nullptr || v == nullptr || B <= 0 || T <= 0 || heads <= 0) return FORDE_EINVAL;
  if (T > AT_ROWS || ld < static_cast<long long>(heads) * AT_DH) return FORDE_EINVAL;
  const uint64_t w = static_cast<uint64_t>(heads) * AT_DH;
  CUtensorMap* maps[8] = {&p.tmQ, &p.tmK, &p.tmV, &p.tmDO, &p.tmO, &p.tmDQ, &p.tmDK, &p.tmDV};
  const void* ptrs[8] = {q, k, v, d_o, o, dq, dk, dv};
  for (int i = 0; i < 8; ++i) {
    if (ptrs[i] == nullptr) continue;
    const int rc = make_tmap_bf16_3d(maps[i], ptrs[i], w, static_cast<uint64_t>(T),
                                     static_cast<uint64_t>(B), static_cast<uint64_t>(ld),
                                     static_cast<uint64_t>(T) * static_cast<uint64_t>(ld), AT_DH, AT_ROWS);
    if (rc != FORDE_OK) return rc;
  }
  p.B = B; p.T = T; p.heads = heads; p.Tpad = (T + 15) / 16 * 16; p.ld = ld;
  p.scale = 0.125f;  // 1 / sqrt(64)
  return FORDE_OK;
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_attn_max_tokens(void) { return AT_ROWS; }

int forde_attn_fwd(void* stream, const void* q, const void* k, const void* v, void* o, float* lse, int B, int T,
                   int heads, long long ld) {
  if (o == nullptr || lse == nullptr) return FORDE_EINVAL;
  if ((reinterpret_cast<uintptr_t>(o) & 15) != 0 || (ld & 7) != 0) return FORDE_EALIGN;
  AttnParams p = {};
  int rc = fill_attn(p, q, k, v, nullptr, B, T, heads, ld, o);
  if (rc != FORDE_OK) return rc;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_fwd_smem());
  });
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const int items = B * heads;
  const int grid = items < 2 * sms ? items : 2 * sms;   // two CTAs per SM (83 KB of shared memory, 256 TMEM columns each)
  attn_fwd_kernel<<<grid, AT_THREADS, attn_fwd_smem(), static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("attn_fwd_kernel");
}

int forde_attn_bwd(void* stream, const void* q, const void* k, const void* v, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int T, int heads, long long ld) {
  if (d_o == nullptr || lse == nullptr || dq == nullptr || dk == nullptr || dv == nullptr) return FORDE_EINVAL;
  if (((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) != 0 ||
      (ld & 7) != 0)
    return FORDE_EALIGN;
  AttnParams p = {};
  int rc = fill_attn(p, q, k, v, d_o, B, T, heads, ld, nullptr, dq, dk, dv);
  if (rc != FORDE_OK) return rc;
  p.lse = const_cast<float*>(lse);
  p.dq = static_cast<__nv_bfloat16*>(dq);
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_bwd_smem());
  });
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const int items = B * heads;
  attn_bwd_kernel<<<items < sms ? items : sms, AT_THREADS, attn_bwd_smem(), static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("attn_bwd_kernel");
}

}  // extern "C"
