// Self-attention core for MEDIUM sequences (T <= 256, head_dim 64): the vision tower of the FORDE dual encoder
// (README.md:30-34,42-43; ViT-B/16 at 224 px: T = 197 tokens = two blocks of 128 query rows, 208 padded keys).
//
// Why it exists: cuDNN's flash kernels need 108 us (forward) + 349 us (backward, three kernels) per vision layer at
// C2 -- 5.5 ms of a 39 ms train step -- for 310 / 620 MB of traffic and 31 / 76 GFLOP.  A (batch, head) item is small
// enough that ALL of its operands fit in shared memory and all of its score tiles in TMEM, so one persistent CTA per SM
// walks the items with the operand loads of the next item in flight (TMA ring) while this one computes.
//
//   forward  (per item): S_i = Q_i K^T for both 128-row blocks i (N = Tk padded keys, <= 256 TMEM columns each)
//             softmax: warps 0-3 own row block 0, warps 4-7 row block 1, one thread per score row: row max, then
//             p = exp2(s * scale * log2e - max) packed to bf16 and stored to that block's swizzled staging buffer,
//             the TMEM load of the next 32 columns in flight while a chunk is processed
//             O_i = P_i V accumulates over the first 64 columns of S_i's own TMEM region, in two parts: keys [0, 128)
//             as soon as they are staged (under the exponentials of the remaining keys), the rest when the row is done
//             read-out: O_i / row sum -> bf16 rows in the (idle) staging buffer -> one TMA store per block
//             issue order: row block 1 runs half an item behind row block 0 (S_0(n), P_1 V of item n - 1, S_1(n),
//             P_0 V of item n), so one block's exponentials run under the other's max pass and P V / read-out tail
//             What bounds it (ncu, profiles/r2_ncu_full_aux_kernels.md): neither pipe saturates (MUFU 38 %, tensor
//             17 %); with 2.5 warps per scheduler the per-warp dependency chains are exposed (7 cycles per issued
//             instruction: fixed-latency waits 1.9, TMEM / barrier scoreboard 1.5).  Per item and block: max pass
//             1.9 k clocks, exp pass 4.7 k, P V wait + read-out 1.4 k (clock64 trace).
//   backward (per item): tiles (key block j, row block i), j outer.  Per tile:
//             S = Q_i K_j^T, dP = dO_i V_j^T                                   -> TMEM [0,128), [128,256)
//             p = exp2(s * scale * log2e - lse), dS = p (dP - D) scale         -> P, dS (bf16) staged in shared memory
//             dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j               -> TMEM [256,320), [320,384), [384+64 i, ..)
//             D = rowsum(dO * O) is taken from the forward output (needed before the first key block is done).
//             The MMA warp issues tile g+1's S / dP BEFORE tile g's gradient MMAs, the softmax warps keep tile g+1's
//             P / dS in registers until tile g's gradient MMAs have released the staging buffers, and the dV / dK / dQ
//             read-outs are deferred behind the next tile's softmax, so the tensor pipe and the MUFU pipe overlap.
// Operand ring (backward): four 32 KB slots {K_j,V_j} / {Q_i,dO_i}, loaded in the order K0V0, Q0dO0, Q1dO1, K1V1 and
// released in the same order (tile 1, 2, 3, 3), so the next item's loads are issued two tiles before their first use.
// What bounds it: the exponentials (16 MUFU lanes per clock per SM: 128 x 128 / 16 = 1024 clocks per tile) and HBM
// (620 MB per backward layer = 95 us); the tensor work (1280 clocks per tile) hides under both.
// Rows / keys beyond T: TMA zero-fills them (3-D tensor map {columns, tokens, batch}); masked keys get p = 0; padded
// query rows get p = dS = 0 in the backward because they are reduction indices of dV and dK.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "../../include/forde_b200.h"
#include "forde_host.h"
#include "sm100_ptx.cuh"

namespace forde {

constexpr int A2_ROWS = 128;
constexpr int A2_DH = 64;
constexpr int A2_TILE = A2_ROWS * A2_DH * 2;     // 16 KB: 128 rows x 64 bf16, 128-byte swizzle
constexpr int A2_MAXT = 256;
constexpr int A2_THREADS = 320;                  // warps 0-7 softmax / epilogue, warp 8 TMA, warp 9 MMA + TMEM
constexpr int A2_SOFT = 256;
constexpr int A2_RING = 4;                       // backward operand ring slots (2 tiles each); a power of two
constexpr float kA2Log2e = 1.4426950408889634f;
constexpr float kA2Ln2 = 0.6931471805599453f;

struct Attn2Params {
  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ, tmDK, tmDV, tmO;
  __nv_bfloat16* o;            // forward output [B*T][ld]
  const __nv_bfloat16* o_in;   // backward: the forward output
  const __nv_bfloat16* do_in;  // backward: the output gradient (row reads for D; the tiles come through tmDO)
  float* lse;                  // [B][heads][T]
  __nv_bfloat16 *dq, *dk, *dv;
  int B, T, heads, Tk, nb;     // Tk = T rounded up to 16; nb = number of 128-row blocks (1 or 2)
  long long ld;
  float scale;
  long long* dbg;   // optional clock64 trace of CTA 0 (tools/attn_trace.py); NULL in production
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void pair_bar_sync(int id) {   // the two warps that share a TMEM lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
// shared address of the 16-byte chunk holding columns [col, col + 8) of row r in a [tiles][128 rows][64 cols] bf16
// buffer with TMA's 128-byte swizzle (tile = col / 64)
__device__ __forceinline__ uint32_t a2_chunk(uint32_t base, int r, int col) {
  return base + static_cast<uint32_t>(col >> 6) * A2_TILE + static_cast<uint32_t>(r) * 128u +
         ((static_cast<uint32_t>((col & 63) >> 3) ^ static_cast<uint32_t>(r & 7)) << 4);
}
// 32 fp32 accumulator values -> 32 bf16 = 64 contiguous bytes
__device__ __forceinline__ void a2_store_row32(__nv_bfloat16* dst, const uint32_t* v, float mul) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * mul, __uint_as_float(v[8 * j + 1]) * mul);
    q.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul);
    q.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul);
    q.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul);
    d4[j] = q;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// ------------------------------------------------------------------------------------------------ forward
// Warps 0-3 own row block 0, warps 4-7 row block 1: a thread holds ONE whole score row (TMEM lane), so the row max and
// the row sum need no exchange between warps, and the two blocks' latency-bound phases (TMEM reads, staging, read-out)
// run side by side instead of one after the other.  clock64 trace of the first version (all eight warps on one block
// at a time, two warps per row, per-thread row stores of O): 13.1 k cycles per item, of which 4.5 k the exponentials
// -- 169 us per ViT-B/16 layer; this one 8.2 k cycles, 114 us (cuDNN: 108; profiles/r2_attn2_forward.md).
// shared memory: Q [2 row blocks] | K [2 key blocks] | V [2 key blocks] | P staging [2 row blocks][4 column tiles]
//                (tile 0 of a block's staging also carries its O on the way out: TMA store) | barriers
// Q / K are single-buffered: both S products are issued at the start of an item, so the next item's Q / K loads have
// the rest of the item to land.
__device__ __forceinline__ float a2_max32(const uint32_t* v, int c0, int T, float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;      // four chains instead of one of 32 dependent ops
  if (c0 + 32 <= T) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      m0 = fmaxf(m0, __uint_as_float(v[j]));
      m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
      m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
      m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < T) m0 = fmaxf(m0, __uint_as_float(v[j]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// p = exp2(s * sl2 - mxs) for the 32 scores of a chunk (0 for keys beyond T), staged as bf16; returns their sum
__device__ __forceinline__ float a2_exp32_stage(const uint32_t* v, int c0, int T, float sl2, float mxs, uint32_t sP_u,
                                                int r) {
  uint32_t pk[16];
  float s0 = 0.f, s1 = 0.f;
  if (c0 + 32 <= T) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mxs));
      const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mxs));
      s0 += p0;
      s1 += p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + 2 * j;
      const float p0 = (c < T) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -mxs)) : 0.f;
      const float p1 = (c + 1 < T) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -mxs)) : 0.f;
      s0 += p0;
      s1 += p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  }
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8)
    st_shared_v4(a2_chunk(sP_u, r, c0 + j8 * 8), pk[4 * j8], pk[4 * j8 + 1], pk[4 * j8 + 2], pk[4 * j8 + 3]);
  return s0 + s1;
}

__global__ void __launch_bounds__(A2_THREADS, 1) attn2_fwd_kernel(const __grid_constant__ Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * A2_TILE;
  uint8_t* sV = smem + 4 * A2_TILE;
  uint8_t* sP = smem + 6 * A2_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 14 * A2_TILE);
  uint64_t* fullQK = bars;
  uint64_t* emptyQK = bars + 1;
  uint64_t* fullV = bars + 2;
  uint64_t* emptyV = bars + 3;
  uint64_t* bar_s = bars + 4;      // [2] S_i complete
  uint64_t* bar_pa = bars + 6;     // [2] keys [0, 128) of P_i staged (128 arrivals): the first part of P_i V may start
  uint64_t* bar_pb = bars + 8;     // [2] all of P_i staged (128 arrivals); every score of S_i has been read
  uint64_t* bar_o = bars + 10;     // [2] O_i = P_i V complete (also: P_i's staging buffer idle again)
  uint64_t* tfree = bars + 12;     // [2] O_i read out of TMEM (128 arrivals): region i may take the next item's S_i
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == A2_SOFT) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmO);
    mbar_init(fullQK, 1);
    mbar_init(emptyQK, 1);
    mbar_init(fullV, 1);
    mbar_init(emptyV, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_pa[i], 128);
      mbar_init(&bar_pb[i], 128);
      mbar_init(&bar_o[i], 1);
      mbar_init(&tfree[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int items = p.B * p.heads;
  const int T = p.T, Tk = p.Tk, nb = p.nb;

  if (warp == 8) {
    if (lane == 0) {
      int n = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        const int b = it / p.heads, h = it - b * p.heads;
        mbar_wait(emptyQK, (n & 1) ^ 1);
        mbar_expect_tx(fullQK, 2 * nb * A2_TILE);
        for (int i = 0; i < nb; ++i) {
          tma_load_3d(&p.tmQ, fullQK, sQ + i * A2_TILE, h * A2_DH, i * A2_ROWS, b);
          tma_load_3d(&p.tmK, fullQK, sK + i * A2_TILE, h * A2_DH, i * A2_ROWS, b);
        }
        mbar_wait(emptyV, (n & 1) ^ 1);
        mbar_expect_tx(fullV, nb * A2_TILE);
        for (int i = 0; i < nb; ++i) tma_load_3d(&p.tmV, fullV, sV + i * A2_TILE, h * A2_DH, i * A2_ROWS, b);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(A2_ROWS, Tk, false, false);   // Q (K-major) x K (K-major)
      const uint32_t idesc_o = make_idesc_bf16(A2_ROWS, A2_DH, false, true); // P (K-major) x V (MN-major)
      const uint64_t dK = make_smem_desc_sw128(smem_u32(sK), 0, 1024);
      const uint64_t dV = make_smem_desc_sw128(smem_u32(sV), A2_TILE, 1024);
      // Issue order.  Row block 1 runs half an item behind row block 0: per item S_0(n), P_1 V of item n - 1, S_1(n),
      // P_0 V of item n.  The two blocks' softmax groups are then in different phases -- one block's exponentials (MUFU)
      // run under the other's TMEM-read-bound max pass and its P V / read-out tail -- instead of both hitting the same
      // unit at the same time (both blocks of an item issued together: 122 instead of 114 us per ViT-B/16 layer).
      // P_i V goes in two parts: the keys [0, 128) as soon as they are staged -- under the exponentials of the remaining
      // keys --, the rest when the row is complete.  (Runtime loops on purpose: fully unrolled with a per-step
      // predicate, this compiler build produced wrong operand registers for the last steps.)
      const int ks_a = (Tk < 128 ? Tk : 128) / 16, ks_n = Tk / 16;
      auto issue_s = [&](int i, int n) {
        const uint64_t dQ = make_smem_desc_sw128(smem_u32(sQ) + i * A2_TILE, 0, 1024);
        if (n > 0) mbar_wait(&tfree[i], (n - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < A2_DH / 16; ++k)
          umma_bf16(tmem_base + i * 256, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bar_s[i]);
      };
      auto issue_pv = [&](int i, int n, bool first_part) {
        const uint64_t dP = make_smem_desc_sw128(smem_u32(sP) + i * 4 * A2_TILE, 0, 1024);
        mbar_wait(first_part ? &bar_pa[i] : &bar_pb[i], n & 1);
        tc_fence_after();
        const int k0 = first_part ? 0 : ks_a, k1 = first_part ? ks_a : ks_n;
        for (int ks = k0; ks < k1; ++ks)         // reduction over the keys, 16 per step
          umma_bf16(tmem_base + i * 256, dP + (ks >> 2) * (A2_TILE >> 4) + (ks & 3) * 2, dV + 128 * ks, idesc_o,
                    ks > 0 ? 1u : 0u);
        if (!first_part) umma_commit(&bar_o[i]);
      };
      int n = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        mbar_wait(fullQK, n & 1);
        issue_s(0, n);
        if (nb > 1) {
          if (n > 0) {                           // block 1 of the previous item: its V is released here
            issue_pv(1, n - 1, true);
            issue_pv(1, n - 1, false);
            umma_commit(emptyV);
          }
          issue_s(1, n);
        }
        umma_commit(emptyQK);
        mbar_wait(fullV, n & 1);
        issue_pv(0, n, true);
        issue_pv(0, n, false);
        if (nb == 1) umma_commit(emptyV);
      }
      if (nb > 1 && n > 0) {                     // the last item's block 1
        issue_pv(1, n - 1, true);
        issue_pv(1, n - 1, false);
      }
    }
  } else if ((warp >> 2) < nb) {
    const int i = warp >> 2, q = warp & 3;                             // row block, TMEM lane quarter
    const int r = q * 32 + lane;                                       // row within the block = TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + i * 256;
    const uint32_t sP_u = smem_u32(sP) + i * 4 * A2_TILE;
    const float sl2 = p.scale * kA2Log2e;
    const int nch = (Tk + 31) >> 5;                                    // 32-column chunks of a score row (<= 8)
    const bool leader = (q == 0 && lane == 0);                         // issues this block's O store
    const int rg = i * A2_ROWS + r;
    const bool trace = p.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 127) == 0;
#define A2F_TR(slot) do { if (trace && n < 16) p.dbg[n * 16 + i * 8 + (slot)] = clock64(); } while (0)
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int b = it / p.heads, h = it - b * p.heads;
      A2F_TR(0);
      mbar_wait(&bar_s[i], n & 1);
      tc_fence_after();
      A2F_TR(1);
      // both passes: the TMEM load of chunk c + 1 is in flight while chunk c is processed
      uint32_t va[32], vb[32];
      float mx = -INFINITY;
      tmem_ld32(taddr, va);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {
          uint32_t* cur = (c & 1) ? vb : va;
          uint32_t* nxt = (c & 1) ? va : vb;
          if (c + 1 < nch) tmem_ld32(taddr + (c + 1) * 32, nxt);
          mx = a2_max32(cur, c * 32, T, mx);
          if (c + 1 < nch) tmem_wait_ld();
        }
      }
      const float mxs = mx * sl2;
      A2F_TR(2);
      tmem_ld32(taddr, va);
      // this block's staging buffer: P V of the previous item has completed (every thread waited for it before its
      // read-out); the TMA store of that item's O must have read tile 0
      if (n > 0) {
        if (leader) tma_wait_read<0>();
        named_bar_sync(1 + i, 128);
      }
      tmem_wait_ld();
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {
          uint32_t* cur = (c & 1) ? vb : va;
          uint32_t* nxt = (c & 1) ? va : vb;
          if (c + 1 < nch) tmem_ld32(taddr + (c + 1) * 32, nxt);
          sum += a2_exp32_stage(cur, c * 32, T, sl2, mxs, sP_u, r);
          if (c == 3 && nch > 4) {               // keys [0, 128) staged: the first part of P V may start
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bar_pa[i]);
          }
          if (c + 1 < nch) tmem_wait_ld();
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      if (nch <= 4) mbar_arrive(&bar_pa[i]);
      mbar_arrive(&bar_pb[i]);
      A2F_TR(3);
      const float inv = 1.f / sum;
      if (rg < T) p.lse[(static_cast<size_t>(b) * p.heads + h) * T + rg] = (mxs + log2f(sum)) * kA2Ln2;
      // read-out: O_i -> bf16 rows in tile 0 of the (idle) staging buffer -> one TMA store, rows beyond T clipped
      mbar_wait(&bar_o[i], n & 1);
      tc_fence_after();
      A2F_TR(4);
      {
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_ld32(taddr + 32, vb);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&tfree[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          st_shared_v4(a2_chunk(sP_u, r, 8 * k),
                       pack_bf16x2(__uint_as_float(va[8 * k]) * inv, __uint_as_float(va[8 * k + 1]) * inv),
                       pack_bf16x2(__uint_as_float(va[8 * k + 2]) * inv, __uint_as_float(va[8 * k + 3]) * inv),
                       pack_bf16x2(__uint_as_float(va[8 * k + 4]) * inv, __uint_as_float(va[8 * k + 5]) * inv),
                       pack_bf16x2(__uint_as_float(va[8 * k + 6]) * inv, __uint_as_float(va[8 * k + 7]) * inv));
          st_shared_v4(a2_chunk(sP_u, r, 32 + 8 * k),
                       pack_bf16x2(__uint_as_float(vb[8 * k]) * inv, __uint_as_float(vb[8 * k + 1]) * inv),
                       pack_bf16x2(__uint_as_float(vb[8 * k + 2]) * inv, __uint_as_float(vb[8 * k + 3]) * inv),
                       pack_bf16x2(__uint_as_float(vb[8 * k + 4]) * inv, __uint_as_float(vb[8 * k + 5]) * inv),
                       pack_bf16x2(__uint_as_float(vb[8 * k + 6]) * inv, __uint_as_float(vb[8 * k + 7]) * inv));
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + i, 128);
      if (leader) {
        tma_store_3d(&p.tmO, sP_u, h * A2_DH, i * A2_ROWS, b);
        tma_commit();
      }
      A2F_TR(5);
    }
#undef A2F_TR
    if (leader) tma_wait_all<0>();               // the last store has landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// shared memory: operand ring [4 slots][2 tiles] | P staging [2 column tiles] | dS staging [2 column tiles] |
//                output staging [2 tiles] (dV / dK, dQ_0 / dQ_1 -> TMA stores) | barriers
// What the clock64 trace of the first version showed (tools/attn_trace.py, 37 k cycles per item): every thread writing
// its own 128-byte row of dV / dK / dQ (32 lines per store instruction) took 3-5 k cycles per read-out, the same
// pattern reading O for D = rowsum(dO * O) 4-5 k cycles per row block, and the MMA thread needed ~80 cycles per MMA
// (descriptors rebuilt with dependent integer chains).  Hence: read-outs staged in shared memory and written by TMA,
// O read 4 rows x 128 contiguous bytes per warp instruction one tile ahead of its use, descriptors built once per
// tile and advanced by adding constants, tile indices kept as counters.
__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  float d = 0.f;
  d = fmaf(bf16lo_to_f32(a.x), bf16lo_to_f32(b.x), d); d = fmaf(bf16hi_to_f32(a.x), bf16hi_to_f32(b.x), d);
  d = fmaf(bf16lo_to_f32(a.y), bf16lo_to_f32(b.y), d); d = fmaf(bf16hi_to_f32(a.y), bf16hi_to_f32(b.y), d);
  d = fmaf(bf16lo_to_f32(a.z), bf16lo_to_f32(b.z), d); d = fmaf(bf16hi_to_f32(a.z), bf16hi_to_f32(b.z), d);
  d = fmaf(bf16lo_to_f32(a.w), bf16lo_to_f32(b.w), d); d = fmaf(bf16hi_to_f32(a.w), bf16hi_to_f32(b.w), d);
  return d;
}

constexpr int A2B_THREADS = 448;   // warps 0-7 softmax, 8-11 read-out / D warps, 12 TMA, 13 MMA + TMEM

__global__ void __launch_bounds__(A2B_THREADS, 1) attn2_bwd_kernel(const __grid_constant__ Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* sP = smem + 2 * A2_RING * A2_TILE;
  uint8_t* sDS = sP + 2 * A2_TILE;
  uint8_t* sOut = sDS + 2 * A2_TILE;
  float* dbuf = reinterpret_cast<float*>(sOut + 2 * A2_TILE);   // [2 row blocks][128] D * scale of the current item
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * A2_TILE + 1024);
  uint64_t* full = bars;                 // [4]
  uint64_t* empty = bars + A2_RING;      // [4]
  uint64_t* bar_sdp = bars + 8;          // S and dP of a tile complete
  uint64_t* bar_pds = bars + 9;          // P and dS of a tile staged (256 arrivals); S / dP columns consumed
  uint64_t* bar_g = bars + 10;           // gradient MMAs of a tile complete (staging buffers free again)
  uint64_t* bar_kv = bars + 11;          // dV_j / dK_j complete (after the tile (j, last row block))
  uint64_t* bar_q = bars + 12;           // dQ_0.. of an item complete (after its last tile)
  uint64_t* kvfree = bars + 13;          // dV_j / dK_j read out of TMEM (128 arrivals)
  uint64_t* qfree = bars + 14;           // dQ_i of an item read out of TMEM (128 arrivals)
  uint64_t* dready = bars + 15;          // [2] D of row block i written to dbuf (128 arrivals)
  uint64_t* dfree = bars + 17;           // [2] D of row block i read by the softmax threads (256 arrivals)
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 19);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 12 * 32) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmDO);
    for (int i = 0; i < A2_RING; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, A2_SOFT);
    mbar_init(bar_g, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(kvfree, 128);
    mbar_init(qfree, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dready[i], 128);
      mbar_init(&dfree[i], A2_SOFT);
    }
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(smem_u32(tmem_ptr_s)));
  const int items = p.B * p.heads;
  const int T = p.T, Tk = p.Tk, nb = p.nb;
  const int lpi = 2 * nb;              // ring loads per item: K0V0, Q0dO0 [, Q1dO1, K1V1]
  const int tpi = nb * nb;             // tiles per item, ordered (j, i) with j outer
  const int my_items = (items > static_cast<int>(blockIdx.x))
                           ? (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;
  const int G = my_items * tpi;

  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int L = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = it / p.heads, h = it - b * p.heads;
        // (The ring holds one item, so the next item's loads are issued only a tile or two before their use.  Pulling its
        // tiles into L2 a whole item ahead with cp.async.bulk.prefetch removed those waits but MEASURED SLOWER -- 218 ->
        // 234 us per layer, same box, alternating runs -- and was deleted.)
        for (int l = 0; l < lpi; ++l, ++L) {
          const int slot = L & (A2_RING - 1);
          uint8_t* dst = ring + slot * 2 * A2_TILE;
          mbar_wait(&empty[slot], ((L / A2_RING) & 1) ^ 1);
          mbar_expect_tx(&full[slot], 2 * A2_TILE);
          if (l == 0 || l == 3) {          // K_j, V_j
            const int j = l == 0 ? 0 : 1;
            tma_load_3d(&p.tmK, &full[slot], dst, h * A2_DH, j * A2_ROWS, b);
            tma_load_3d(&p.tmV, &full[slot], dst + A2_TILE, h * A2_DH, j * A2_ROWS, b);
          } else {                         // Q_i, dO_i
            const int i = l - 1;
            tma_load_3d(&p.tmQ, &full[slot], dst, h * A2_DH, i * A2_ROWS, b);
            tma_load_3d(&p.tmDO, &full[slot], dst + A2_TILE, h * A2_DH, i * A2_ROWS, b);
          }
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_tt = make_idesc_bf16(A2_ROWS, A2_DH, true, true);    // P^T x dO, dS^T x Q (both MN-major)
      const uint32_t idesc_kt = make_idesc_bf16(A2_ROWS, A2_DH, false, true);   // dS (K-major) x K (MN-major)
      const uint32_t ring_u = smem_u32(ring);
      // descriptors of the staging buffers never change; operand descriptors are rebuilt per tile from the slot base
      const uint64_t dP_mn = make_smem_desc_sw128(smem_u32(sP), A2_TILE, 1024);     // P^T  (M = keys), 16 rows / step
      const uint64_t dDS_mn = make_smem_desc_sw128(smem_u32(sDS), A2_TILE, 1024);   // dS^T (M = keys)
      const uint64_t dDS_k = make_smem_desc_sw128(smem_u32(sDS), 0, 1024);          // dS   (K-major), 16 keys / step
      const uint64_t d_k0 = make_smem_desc_sw128(0, 0, 1024);                       // templates: add (address >> 4)
      const uint64_t d_mn0 = make_smem_desc_sw128(0, A2_TILE, 1024);
      auto issue_sdp = [&](int Lkv, int Lq, int j) {
        mbar_wait(&full[Lkv & (A2_RING - 1)], (Lkv / A2_RING) & 1);
        mbar_wait(&full[Lq & (A2_RING - 1)], (Lq / A2_RING) & 1);
        tc_fence_after();
        const uint32_t kv16 = (ring_u + (Lkv & (A2_RING - 1)) * 2 * A2_TILE) >> 4;
        const uint32_t q16 = (ring_u + (Lq & (A2_RING - 1)) * 2 * A2_TILE) >> 4;
        const uint64_t dQ = d_k0 + q16, dDO = d_k0 + q16 + (A2_TILE >> 4);
        const uint64_t dK = d_k0 + kv16, dV = d_k0 + kv16 + (A2_TILE >> 4);
        const int nk = min(A2_ROWS, Tk - j * A2_ROWS);
        const uint32_t idesc_s = make_idesc_bf16(A2_ROWS, nk, false, false);
#pragma unroll
        for (int k = 0; k < A2_DH / 16; ++k)      // S = Q_i K_j^T
          umma_bf16(tmem_base, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < A2_DH / 16; ++k)      // dP = dO_i V_j^T
          umma_bf16(tmem_base + 128, dDO + 2 * k, dV + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_sdp);
      };
      int nl = 0, j = 0, i = 0;                   // tile g = (item nl of this CTA, key block j, row block i)
      if (G > 0) issue_sdp(0, 1, 0);
      for (int g = 0; g < G; ++g) {
        int nnl = nl, nj = j, ni = i + 1;         // tile g + 1
        if (ni == nb) { ni = 0; if (++nj == nb) { nj = 0; ++nnl; } }
        const int Lkv = nl * lpi + (j == 0 ? 0 : 3), Lq = nl * lpi + 1 + i;
        mbar_wait(bar_pds, g & 1);                // P / dS of tile g staged, its S / dP columns consumed
        if (g + 1 < G)                            // the next tile's scores first: its softmax overlaps the MMAs below
          issue_sdp(nnl * lpi + (nj == 0 ? 0 : 3), nnl * lpi + 1 + ni, nj);
        const int kvgroup = nl * nb + j;
        if (i == 0 && kvgroup > 0) mbar_wait(kvfree, (kvgroup - 1) & 1);   // dV / dK columns were read out
        if (j == 0 && nl > 0) mbar_wait(qfree, (nl - 1) & 1);              // dQ columns were read out
        tc_fence_after();
        const uint32_t kv16 = (ring_u + (Lkv & (A2_RING - 1)) * 2 * A2_TILE) >> 4;
        const uint32_t q16 = (ring_u + (Lq & (A2_RING - 1)) * 2 * A2_TILE) >> 4;
        const uint64_t dQ_mn = d_mn0 + q16, dDO_mn = d_mn0 + q16 + (A2_TILE >> 4), dK_mn = d_mn0 + kv16;
        const int nk = min(A2_ROWS, Tk - j * A2_ROWS);
#pragma unroll
        for (int ks = 0; ks < A2_ROWS / 16; ++ks) {   // reductions over the 128 query rows of block i (2048 B / step)
          const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
          umma_bf16(tmem_base + 256, dP_mn + 128 * ks, dDO_mn + 128 * ks, idesc_tt, acc);   // dV_j += P^T dO_i
          umma_bf16(tmem_base + 320, dDS_mn + 128 * ks, dQ_mn + 128 * ks, idesc_tt, acc);   // dK_j += dS^T Q_i
        }
        for (int ks = 0; ks < nk / 16; ++ks)          // dQ_i += dS K_j over the keys of block j
          umma_bf16(tmem_base + 384 + 64 * i, dDS_k + (ks >> 2) * (A2_TILE >> 4) + (ks & 3) * 2, dK_mn + 128 * ks,
                    idesc_kt, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_g);
        if (i == nb - 1) {
          umma_commit(bar_kv);                          // dV_j, dK_j are final: the read-out warps take them
          umma_commit(&empty[Lkv & (A2_RING - 1)]);     // last use of K_j, V_j
        }
        if (j == nb - 1) umma_commit(&empty[Lq & (A2_RING - 1)]);    // last use of Q_i, dO_i
        if (i == nb - 1 && j == nb - 1) umma_commit(bar_q);           // dQ of the item is final
        nl = nnl; j = nj; i = ni;
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ read-out / D warps (one per TMEM lane quarter)
    // They take everything off the softmax warps' critical path: D = rowsum(dO * O) of the NEXT row blocks (O read
    // 4 rows x 128 contiguous bytes per warp instruction) and the dV / dK / dQ read-outs (TMEM -> bf16 rows in a
    // swizzled staging tile -> one TMA store each, rows beyond T clipped by the tensor map).
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t ring_u = smem_u32(ring), sOut_u = smem_u32(sOut), dbuf_u = smem_u32(dbuf);
    const uint32_t row_sw = static_cast<uint32_t>(r) * 128u, rx = static_cast<uint32_t>(r & 7);
    const bool leader = (warp == 8 && lane == 0);      // issues the TMA stores
    const int dm = lane >> 3, dc = lane & 7;
    int stage = 0;                                     // staging tiles alternate

    auto read_out = [&](uint32_t tcol, const CUtensorMap* tm, int c0, int c1, int c2) {
      if (leader) tma_wait_read<1>();                  // the store that used this staging tile two read-outs ago was read
      named_bar_sync(5, 128);
      const uint32_t dst = sOut_u + stage * A2_TILE + row_sw;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32];
        tmem_ld32(taddr + tcol + hf * 32, a);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          st_shared_v4(dst + ((static_cast<uint32_t>(hf * 4 + k) ^ rx) << 4),
                       pack_bf16x2(__uint_as_float(a[8 * k]), __uint_as_float(a[8 * k + 1])),
                       pack_bf16x2(__uint_as_float(a[8 * k + 2]), __uint_as_float(a[8 * k + 3])),
                       pack_bf16x2(__uint_as_float(a[8 * k + 4]), __uint_as_float(a[8 * k + 5])),
                       pack_bf16x2(__uint_as_float(a[8 * k + 6]), __uint_as_float(a[8 * k + 7])));
      }
      fence_proxy_async_smem();
      named_bar_sync(5, 128);
      if (leader) {
        tma_store_3d(tm, sOut_u + stage * A2_TILE, c0, c1, c2);
        tma_commit();
      }
      stage ^= 1;
    };
    // D of row block ii of `item` (the n-th item of this CTA) -> dbuf.  Both operands are read from global memory (dO a
    // second time, an L2 hit next to the TMA load of the same tile) so that D does not wait for the operand ring: it is
    // computed a whole item ahead of its use.
    auto compute_d = [&](int item, int n, int ii) {
      const int b = item / p.heads, h = item - b * p.heads;
      uint4 o4[8], g4[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rg = ii * A2_ROWS + q * 32 + 4 * k + dm;
        o4[k] = make_uint4(0u, 0u, 0u, 0u);
        g4[k] = make_uint4(0u, 0u, 0u, 0u);
        if (rg < T) {
          const size_t off = (static_cast<size_t>(b) * T + rg) * p.ld + h * A2_DH;
          o4[k] = __ldg(reinterpret_cast<const uint4*>(p.o_in + off) + dc);
          g4[k] = __ldg(reinterpret_cast<const uint4*>(p.do_in + off) + dc);
        }
      }
      float D = 0.f, part[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) part[k] = dot8_bf16(g4[k], o4[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 2);
#pragma unroll
      for (int k = 0; k < 8; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 4);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = __shfl_sync(0xffffffffu, part[k], (lane & 3) * 8);   // row 4 k + (lane & 3) of this quarter
        if ((lane >> 2) == k) D = v;
      }
      if (n > 0) mbar_wait(&dfree[ii], (n - 1) & 1);   // the softmax threads have taken the previous item's D
      st_shared_f32(dbuf_u + (ii * A2_ROWS + r) * 4, D * p.scale);
      mbar_arrive(&dready[ii]);
    };

    for (int ii = 0; ii < nb && my_items > 0; ++ii) compute_d(blockIdx.x, 0, ii);
    for (int n = 0; n < my_items; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / p.heads, h = item - b * p.heads;
      if (n + 1 < my_items)                            // the next item's D, a whole item ahead
        for (int ii = 0; ii < nb; ++ii) compute_d(item + gridDim.x, n + 1, ii);
      for (int jj = 0; jj < nb; ++jj) {
        mbar_wait(bar_kv, (n * nb + jj) & 1);
        tc_fence_after();
        read_out(256, &p.tmDV, h * A2_DH, jj * A2_ROWS, b);
        read_out(320, &p.tmDK, h * A2_DH, jj * A2_ROWS, b);
        tc_fence_before();
        mbar_arrive(kvfree);
      }
      mbar_wait(bar_q, n & 1);
      tc_fence_after();
      for (int ii = 0; ii < nb; ++ii) read_out(384 + 64 * ii, &p.tmDQ, h * A2_DH, ii * A2_ROWS, b);
      tc_fence_before();
      mbar_arrive(qfree);
    }
    if (leader) tma_wait_all<0>();                     // the last stores have landed before the CTA retires
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int q = warp & 3, ch = warp >> 2;
    const int r = q * 32 + lane;                                       // row within a 128-row block = TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t sP_u = smem_u32(sP), sDS_u = smem_u32(sDS), dbuf_u = smem_u32(dbuf);
    const uint32_t row_sw = static_cast<uint32_t>(r) * 128u, rx = static_cast<uint32_t>(r & 7);
    const float sl2 = p.scale * kA2Log2e;
    float lse2[2] = {0.f, 0.f}, dsc[2] = {0.f, 0.f};                   // per row block: lse * log2e, D * scale
    const bool trace = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define A2_TR(slot) do { if (trace && g < 16) p.dbg[g * 16 + (slot)] = clock64(); } while (0)
    auto load_lse = [&](int item, int ii) -> float {                   // padded rows: +inf, so p = exp2(-inf) = 0
      const int b = item / p.heads, h = item - b * p.heads;
      const int rg = ii * A2_ROWS + r;
      return (rg < T) ? p.lse[(static_cast<size_t>(b) * p.heads + h) * T + rg] : INFINITY;
    };
    int nl = 0, j = 0, i = 0;                          // tile g = (item nl of this CTA, key block j, row block i)
    float lpre = G > 0 ? load_lse(blockIdx.x, 0) : 0.f;
    for (int g = 0; g < G; ++g) {
      A2_TR(0);
      int nnl = nl, nj = j, ni = i + 1;                // tile g + 1
      if (ni == nb) { ni = 0; if (++nj == nb) { nj = 0; ++nnl; } }
      if (j == 0) {                                    // this row block's D (from the read-out warps) and log-sum-exp
        mbar_wait(&dready[i], nl & 1);
        const float d = ld_shared_f32(dbuf_u + (i * A2_ROWS + r) * 4);
        mbar_arrive(&dfree[i]);
        if (i == 0) { lse2[0] = lpre * kA2Log2e; dsc[0] = d; } else { lse2[1] = lpre * kA2Log2e; dsc[1] = d; }
      }
      if (g + 1 < G && nj == 0) lpre = load_lse(blockIdx.x + nnl * gridDim.x, ni);   // used a whole tile from now
      const float l2 = i == 0 ? lse2[0] : lse2[1];
      const float dn = i == 0 ? -dsc[0] : -dsc[1];
      const int nk = min(A2_ROWS, Tk - j * A2_ROWS);
      A2_TR(1);
      mbar_wait(bar_sdp, g & 1);
      tc_fence_after();
      A2_TR(2);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = ch * 64 + cc * 32;
        if (c0 < nk) {
          uint32_t pw[16], dw[16];
          const int cg = j * A2_ROWS + c0;                             // global key index of the chunk's first column
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                             // 16 columns at a time (register budget)
            uint32_t s[16], d[16];
            tmem_ld16(taddr + c0 + hh * 16, s);
            tmem_ld16(taddr + 128 + c0 + hh * 16, d);
            tmem_wait_ld();
            if (cg + 32 <= T) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * k]), sl2, -l2));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * k + 1]), sl2, -l2));
                pw[hh * 8 + k] = pack_bf16x2(p0, p1);
                dw[hh * 8 + k] = pack_bf16x2(p0 * fmaf(__uint_as_float(d[2 * k]), p.scale, dn),
                                             p1 * fmaf(__uint_as_float(d[2 * k + 1]), p.scale, dn));
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int c = cg + hh * 16 + 2 * k;
                const float p0 = (c < T) ? ex2_approx(fmaf(__uint_as_float(s[2 * k]), sl2, -l2)) : 0.f;
                const float p1 = (c + 1 < T) ? ex2_approx(fmaf(__uint_as_float(s[2 * k + 1]), sl2, -l2)) : 0.f;
                pw[hh * 8 + k] = pack_bf16x2(p0, p1);
                dw[hh * 8 + k] = pack_bf16x2(p0 * fmaf(__uint_as_float(d[2 * k]), p.scale, dn),
                                             p1 * fmaf(__uint_as_float(d[2 * k + 1]), p.scale, dn));
              }
            }
          }
          // the staging buffers are free once tile g - 1's gradient MMAs have completed: by the time the first chunk is
          // computed that is the usual case, and the stores of chunk 0 overlap the exponentials of chunk 1
          if (cc == 0) {
            A2_TR(3);
            if (g > 0) mbar_wait(bar_g, (g - 1) & 1);
            A2_TR(4);
          }
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const uint32_t c8 = static_cast<uint32_t>((c0 >> 3) + j8);
            const uint32_t off = (c8 >> 3) * A2_TILE + row_sw + (((c8 & 7u) ^ rx) << 4);
            st_shared_v4(sP_u + off, pw[4 * j8], pw[4 * j8 + 1], pw[4 * j8 + 2], pw[4 * j8 + 3]);
            st_shared_v4(sDS_u + off, dw[4 * j8], dw[4 * j8 + 1], dw[4 * j8 + 2], dw[4 * j8 + 3]);
          }
        } else if (cc == 0) {
          if (g > 0) mbar_wait(bar_g, (g - 1) & 1);                   // (nothing of this tile in my columns)
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
      A2_TR(5);
      nl = nnl; j = nj; i = ni;
    }
#undef A2_TR
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr int attn2_fwd_smem() { return 14 * A2_TILE + 128 + 1024; }
constexpr int attn2_bwd_smem() { return (2 * A2_RING + 6) * A2_TILE + 1024 + 256 + 1024; }

static int fill_attn2(Attn2Params& p, const void* q, const void* k, const void* v, const void* d_o, int B, int T,
                      int heads, long long ld, void* dq = nullptr, void* dk = nullptr, void* dv = nullptr) {
  if (q == nullptr || k == nullptr || v == nullptr || B <= 0 || T <= 0 || heads <= 0) return FORDE_EINVAL;
  if (T > A2_MAXT || ld < static_cast<long long>(heads) * A2_DH) return FORDE_EINVAL;
  const uint64_t w = static_cast<uint64_t>(heads) * A2_DH;
  CUtensorMap* maps[7] = {&p.tmQ, &p.tmK, &p.tmV, &p.tmDO, &p.tmDQ, &p.tmDK, &p.tmDV};
  const void* ptrs[7] = {q, k, v, d_o, dq, dk, dv};
  for (int i = 0; i < 7; ++i) {
    if (ptrs[i] == nullptr) continue;
    const int rc = make_tmap_bf16_3d(maps[i], ptrs[i], w, static_cast<uint64_t>(T), static_cast<uint64_t>(B),
                                     static_cast<uint64_t>(ld), static_cast<uint64_t>(T) * static_cast<uint64_t>(ld),
                                     A2_DH, A2_ROWS);
    if (rc != FORDE_OK) return rc;
  }
  p.B = B; p.T = T; p.heads = heads; p.Tk = (T + 15) / 16 * 16; p.nb = (T + A2_ROWS - 1) / A2_ROWS; p.ld = ld;
  p.scale = 0.125f;  // 1 / sqrt(64)
  return FORDE_OK;
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_attn2_max_tokens(void) { return A2_MAXT; }

// debug: clock64 timeline of CTA 0's first softmax thread (tools/attn_trace.py); NULL in production
static long long* g_attn2_dbg = nullptr;
void forde_attn2_set_trace(void* buf) { g_attn2_dbg = static_cast<long long*>(buf); }

int forde_attn2_fwd(void* stream, const void* q, const void* k, const void* v, void* o, float* lse, int B, int T,
                    int heads, long long ld) {
  if (o == nullptr || lse == nullptr) return FORDE_EINVAL;
  if ((reinterpret_cast<uintptr_t>(o) & 15) != 0 || (ld & 7) != 0) return FORDE_EALIGN;
  Attn2Params p = {};
  int rc = fill_attn2(p, q, k, v, nullptr, B, T, heads, ld);
  if (rc != FORDE_OK) return rc;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.dbg = g_attn2_dbg;
  rc = make_tmap_bf16_3d(&p.tmO, o, static_cast<uint64_t>(heads) * A2_DH, static_cast<uint64_t>(T),
                         static_cast<uint64_t>(B), static_cast<uint64_t>(ld),
                         static_cast<uint64_t>(T) * static_cast<uint64_t>(ld), A2_DH, A2_ROWS);
  if (rc != FORDE_OK) return rc;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn2_fwd_smem());
  });
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const int items = B * heads;
  attn2_fwd_kernel<<<items < sms ? items : sms, A2_THREADS, attn2_fwd_smem(), static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("attn2_fwd_kernel");
}

int forde_attn2_bwd(void* stream, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                    const float* lse, void* dq, void* dk, void* dv, int B, int T, int heads, long long ld) {
  if (o == nullptr || d_o == nullptr || lse == nullptr || dq == nullptr || dk == nullptr || dv == nullptr)
    return FORDE_EINVAL;
  if (((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) |
        reinterpret_cast<uintptr_t>(o)) & 15) != 0 || (ld & 7) != 0)
    return FORDE_EALIGN;
  Attn2Params p = {};
  int rc = fill_attn2(p, q, k, v, d_o, B, T, heads, ld, dq, dk, dv);
  if (rc != FORDE_OK) return rc;
  p.o_in = static_cast<const __nv_bfloat16*>(o);
  p.do_in = static_cast<const __nv_bfloat16*>(d_o);
  p.dbg = g_attn2_dbg;
  p.lse = const_cast<float*>(lse);
  p.dq = static_cast<__nv_bfloat16*>(dq);
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(attn2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn2_bwd_smem());
  });
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  const int items = B * heads;
  attn2_bwd_kernel<<<items < sms ? items : sms, A2B_THREADS, attn2_bwd_smem(), static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("attn2_bwd_kernel");
}

}  // extern "C"
