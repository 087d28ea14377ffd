// Persistent warp-specialised tcgen05 GEMM for sm_100a with the FORDE StatefulLayer epilogues.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]        (bf16 operands, fp32 accumulation in TMEM)
//
// Two instantiations of the tile engine:
//   PAIR = true  (default for large problems): a 2-CTA cluster works on one 256 x 256 tile with
//                tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and only HALF of B (128 of the 256
//                columns), the leader CTA's single thread issues the MMA for both SMs, and each CTA's TMEM receives
//                its 128 x 256 half of the accumulator.  Halving the B traffic is what takes the 128 x 256
//                cta_group::1 tile off the SM<->L2 operand-bandwidth limit (576 KB -> 384 KB of loads per
//                128 x 256 x 768 tile).
//   PAIR = false : one CTA per 128 x 256 tile (small problems, where 256-row tiles would leave SMs idle).
//
// Roles (64 + 32 x epilogue-warps threads, one CTA per SM):
//   warp 0     : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer (128 x 256 x 16, or 256 x 256 x 16 for a pair)
//   warps 2..  : epilogue (8 warps as shipped): tcgen05.ld the accumulator (TMEM lane quarter = warp % 4, one row per
//                thread; the warps of a quarter split the tile's 256 columns), apply the fused epilogue, stage
//                through swizzled smem and TMA-store.  Two epilogue warps per SM sub-partition is what hides the
//                tcgen05.ld / MUFU / st.shared latencies of the fused epilogues (one warp per scheduler left the
//                tensor pipe 38 % busy; see profiles/).
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
//
// Tile scheduling.  Static (single-CTA instantiations, grouped MoE modes): one CTA per SM, work item t = worker,
// worker + #workers, ...  Dynamic (CTA-pair instantiations, `clc`): the grid has ONE CLUSTER PER WORK ITEM, a cluster
// processes its own item and then keeps asking the hardware for further, not yet launched clusters of the grid
// (clusterlaunchcontrol.try_cancel, SASS UGETNEXTWORKID) and does their work too.  Resident clusters therefore behave
// like a persistent grid (prologue once per SM), but the schedule adapts to SMs that are not available: when an NCCL
// kernel of the overlapped gradient exchange holds some SMs, the clusters that cannot become resident there simply never
// start and their tiles are taken by the others, instead of delaying a static 1/74 share of the tiles to the end of the
// kernel (measured round 2: +1.4 ms per step at 2 GPUs whatever NCCL's footprint, profiles/r2_n2_*.md).
// The leader CTA's producer thread issues the query for item j+1 when it STARTS loading item j (one item of
// lookahead, so early clusters cannot hoard work); the response is multicast into a 4-slot ring in both CTAs and
// every role (2 producers, MMA issuer, 2 x 8 epilogue warps) consumes every response once, releasing the slot on the
// leader's `clc_empty` barrier.
//
// Either operand may be K-major (reduction index contiguous in global memory) or MN-major (row index
// contiguous); the smem tile layout and the UMMA descriptors are derived from that, so no operand is ever
// transposed in HBM.
//
// Epilogues (reference: StatefulLayer spec, /root/reference/README.md:47-56; sensing.py:34-69):
//   EPI_STORE   : D (+ per-column bias) -> bf16, TMA store.                       (down projection, dx)
//   EPI_STORE_ADD : D (+ bias) + Cin -> bf16: the Cin tile (may be D itself: accumulate) is TMA-prefetched like the
//                 z tile of EPI_DZ_BWD and overwritten in place.  (input gradients that sum several GEMMs)
//   EPI_UP_FWD  : rows are neurons. z = acc + b[row] -> bf16; h = F_a(z) + gate_a * z -> bf16 (always stored);
//                 z^T stored only where the backward needs it (save_z / warps with tanh neurons);
//                 per-neuron sum|z|, sum z, sum z^2 over each 64-sample chunk -> partials[4 nb + c][3][H].
//   EPI_DZ_BWD  : rows are neurons. TMA-prefetches the z^T box (tanh warps) or the h^T box (sign only),
//                 dz = acc * (F_a'(z) + gate_a) (binary_step uses the straight-through estimator) -> bf16, written in
//                 place over that box; per-neuron sum|dz|, sum dz, sum dz^2 -> partials.
//   EPI_F32     : D * row_scale[m] * col_scale[n] * scale -> fp32, plain store or red.add (split-K dW).
//   EPI_MOE_UP  : rows are routed token slots (MoE variant of the slot, moe.py:17-29).  z = acc + b_e[col] -> bf16;
//                 h = gelu_tanh(z) -> bf16.
//   EPI_MOE_DZ  : dz = acc * gelu_tanh'(z) -> bf16, z tile TMA-prefetched, written in place.
//
// Grouped modes for the MoE experts (the rows of A are routed token slots sorted by expert, every expert's
// segment padded to a multiple of 128 rows; all device-side, no host synchronisation on the routing result):
//   grouped-M : the number of 128-row blocks comes from device memory (m_blocks_dev) and each block carries its
//               expert (group_of_mblock); the B operand and the column bias of that expert are selected by
//               coordinate offsets into the expert-stacked weight / bias arrays.
//   grouped-K : one work item per (expert, output tile); the reduction runs over that expert's row segment
//               (seg_offsets) and the fp32 result goes to the expert's slice of the output (dW per expert).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>

#include "sm100_ptx.cuh"

namespace forde {

enum EpiKind : int {
  EPI_STORE = 0, EPI_UP_FWD = 1, EPI_DZ_BWD = 2, EPI_F32 = 3, EPI_MOE_UP = 4, EPI_MOE_DZ = 5, EPI_STORE_ADD = 6
};

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
__host__ __device__ constexpr int b_tile_bytes(bool pair) { return (pair ? BN / 2 : BN) * BK * 2; }  // 16 / 32 KB
__host__ __device__ constexpr int stage_bytes(bool pair) { return A_TILE_BYTES + b_tile_bytes(pair); }
constexpr int CHUNK = 64;                         // epilogue column chunk (one 128-byte bf16 row)
constexpr int BOX_BYTES = 32 * CHUNK * 2;         // one warp's staging box: 32 rows x 64 bf16 = 4 KB
constexpr int TMEM_COLS = 512;
constexpr int CLC_SLOTS = 4;                      // ring of cluster-launch-control responses (dynamic scheduling)
// Epilogue warps per CTA: 8 = 2 per SM sub-partition, each draining two adjacent 64-column chunks of a tile; the fused
// kernels keep a 5-stage operand ring next to their staging boxes.  16 warps / 3 stages was measured twice and lost
// both times: at C2 vision dims once the epilogues were warp-uniform (dz_bwd 221 vs 192 us, profiles/r1_exp11) and at
// the text-tower dims in round 2 (K = 512: up_fwd 57.3 vs 43.6 us, dz_bwd 51.8 vs 46.2 us, profiles/r2_w16_text_dims.txt)
// -- the main loop starves with 3 stages and the extra warps do not shorten the per-tile store chain.
constexpr int FUSED_EPI_WARPS = 8;
constexpr int FUSED_STAGES = 5;
// Software pipelining of the accumulator reads (tcgen05.ld) in the fused epilogues: 0 = load a 32-column half, use
// it, load the next; 1 = the chunk's second half is in flight while the first is processed; 2 = additionally the next
// chunk's first half is requested before the second half is processed (64 more live registers, still no spills).
// Measured at C2 vision dims (profiles/r1_exp12_tmem_pipe.txt): up_fwd 164.9 -> 161.9 us (all-generalist map),
// 215.6 -> 210.8 us (i.i.d. map) with 2; dz_bwd unchanged within noise -- the epilogues are not bound by the TMEM
// read latency.  -D FORDE_TMEM_PIPE=n sets both for A/B builds.
#ifdef FORDE_TMEM_PIPE
#define FORDE_TMEM_PIPE_UP FORDE_TMEM_PIPE
#define FORDE_TMEM_PIPE_DZ FORDE_TMEM_PIPE
#endif
#ifndef FORDE_TMEM_PIPE_UP
#define FORDE_TMEM_PIPE_UP 2
#endif
#ifndef FORDE_TMEM_PIPE_DZ
#define FORDE_TMEM_PIPE_DZ 0
#endif
// Forward epilogue arithmetic and the sign-bit buffer (adopted after the round-2 A/B, profiles/r2_ab1_sign_bits.txt:
// up_fwd 163.8 us / dz_bwd 162.9 us against 163.2 / 191.0 us for the h^T-reading backward at C2 vision dims):
//  * packed bf16x2 arithmetic: h = z + max(z, 0) of the all-relu warp mode is HMNMX2 + HFMA2 on the packed word and
//    the sign bits come from one HSET2 per pair.  The packed compare yields the pair's two bits 16 apart, so a sign
//    word is INTERLEAVED: within a 32-sample word, sample 2k -> bit k, sample 2k + 1 -> bit 16 + k.
//  * the sign buffer is word-major, [sign_ld words][M neurons]: a warp's 32 lanes are 32 consecutive neurons, so each
//    of its sign stores and loads touches one 128-byte line (the neuron-major layout cost the forward ~12 us: 32
//    lines a row pitch apart per store instruction, +50 % L2 write transactions for 6 % of the bytes).
// forde_sl_sign_layout() reports both properties (= 3) to callers that inspect the bits.
__device__ __forceinline__ long long sign_index(long long row, long long word, int rows) { return word * rows + row; }
__host__ __device__ constexpr int tmem_pipe(int epi) {
  return epi == 1 /*EPI_UP_FWD*/ ? FORDE_TMEM_PIPE_UP : epi == 2 /*EPI_DZ_BWD*/ ? FORDE_TMEM_PIPE_DZ : 0;
}
// Per-neuron partial-sum slots per 256-column tile: each epilogue warp drains two adjacent 64-column chunks and sums
// them into ONE slot (2 slots of 128 samples per tile).
constexpr int PARTIALS_PER_TILE = 2;
constexpr int CHUNKS_PER_SLOT = (BN / CHUNK) / PARTIALS_PER_TILE;
__host__ __device__ constexpr int epi_warps(int epi, bool pair) {
  return FUSED_EPI_WARPS;
}
__host__ __device__ constexpr int num_threads(int epi, bool pair) { return 64 + 32 * epi_warps(epi, pair); }


// EPI_DZ_BWD keeps a per-warp ring of z^T boxes: chunk g is TMA-loaded DZ_DIST chunks ahead into box g % ring,
// multiplied in place into dz and TMA-stored from there.  ring - DZ_DIST - 1 stores may still be reading smem
// when a box is refilled, so with ring 4 / distance 2 the refill never waits on the store just issued.
__host__ __device__ constexpr bool epi_z_ring(int epi) {
  return epi == EPI_DZ_BWD || epi == EPI_MOE_DZ || epi == EPI_STORE_ADD;
}
__host__ __device__ constexpr bool epi_two_out(int epi) { return epi == EPI_UP_FWD || epi == EPI_MOE_UP; }
__host__ __device__ constexpr bool epi_col_bias(int epi) {
  return epi == EPI_STORE || epi == EPI_MOE_UP || epi == EPI_STORE_ADD;
}
__host__ __device__ constexpr bool epi_neuron_rows(int epi) { return epi == EPI_UP_FWD || epi == EPI_DZ_BWD; }
__host__ __device__ constexpr int dz_ring(bool pair) { return 2; }
__host__ __device__ constexpr int dz_dist(bool pair) { return 1; }
// per-warp staging: EPI_STORE [out] ; EPI_UP_FWD [z][h] ; EPI_DZ_BWD [ring of zin/out boxes]
__host__ __device__ constexpr int epi_warp_staging_bytes(int epi, bool pair) {
  return epi == EPI_F32 ? 0 : epi == EPI_STORE ? BOX_BYTES : epi_z_ring(epi) ? dz_ring(pair) * BOX_BYTES : 2 * BOX_BYTES;
}

__host__ __device__ constexpr int gemm_stages(int epi, bool pair) {
  return epi == EPI_UP_FWD ? (pair ? FUSED_STAGES : 3)
         : epi == EPI_DZ_BWD ? (pair ? FUSED_STAGES : 3)
         : (epi == EPI_MOE_UP || epi == EPI_MOE_DZ) ? 3
         : epi == EPI_STORE_ADD                     ? (pair ? 5 : 3)
                                                    : (pair ? 6 : 4);
}
__host__ __device__ constexpr int gemm_smem_bytes(int epi, bool pair) {
  // ring + staging + column bias (BN fp32) + barriers/pointers (512 B) + CLC responses (64 B, padded to 128)
  // + 1 KB alignment slack
  return gemm_stages(epi, pair) * stage_bytes(pair) + epi_warps(epi, pair) * epi_warp_staging_bytes(epi, pair) + BN * 4 + 512 + 128 + 1024;
}
static_assert(gemm_smem_bytes(EPI_STORE, false) <= 232448 && gemm_smem_bytes(EPI_STORE, true) <= 232448, "smem");
static_assert(gemm_smem_bytes(EPI_UP_FWD, false) <= 232448 && gemm_smem_bytes(EPI_UP_FWD, true) <= 232448, "smem");
static_assert(gemm_smem_bytes(EPI_F32, false) <= 232448 && gemm_smem_bytes(EPI_F32, true) <= 232448, "smem");
static_assert(gemm_smem_bytes(EPI_DZ_BWD, false) <= 232448 && gemm_smem_bytes(EPI_DZ_BWD, true) <= 232448, "smem");
static_assert(gemm_smem_bytes(EPI_STORE_ADD, false) <= 232448 && gemm_smem_bytes(EPI_STORE_ADD, true) <= 232448, "smem");

struct GemmParams {
  CUtensorMap tmA;   // operand A   (box 64 x BM for K-major, 64 x 64 for MN-major)
  CUtensorMap tmB;   // operand B
  CUtensorMap tmC0;  // bf16 output 0 (box 64 x 32): y / dx / z^T / dz^T
  CUtensorMap tmC1;  // bf16 output 1 (box 64 x 32): h^T               (EPI_UP_FWD)
  CUtensorMap tmZ;   // bf16 input z^T (box 64 x 32)                   (EPI_DZ_BWD)
  CUtensorMap tmH;   // bf16 input h^T (box 64 x 32): stands in for z^T where no neuron of the warp is tanh (EPI_DZ_BWD)
  int M, N, K;       // D is M x N, reduction length K
  int m_blocks, n_blocks, k_blocks;
  int split_k, kb_per_split;
  int raster_m_fast;  // 1: consecutive work items walk M first (share the B slab), 0: walk N first
  int pair;           // 1: launched as 2-CTA clusters (256-row tiles, cta_group::2)
  int clc;            // 1: one cluster per work item + cluster launch control (pair kernels only), 0: static schedule
  int save_z;         // EPI_UP_FWD: 1 = store every z box; 0 = only the boxes the backward needs (warps with tanh neurons)
  // Optional sign bits of z, word-major [sign_ld words][M rows = neurons] u32; word w of a neuron covers samples
  // 32 w .. 32 w + 31 in the interleaved bit order described at sign_index() (forde_sl_sign_layout() == 3).
  // EPI_UP_FWD writes them for the 32-neuron groups whose z box it does not store; EPI_DZ_BWD then reads 8 bytes
  // per neuron and 64-sample chunk instead of TMA-loading the h^T box (1/16 of the bytes).  Null = off.
  uint32_t* sign_bits;
  int sign_ld;        // row pitch in u32 words: 8 * ceil(N / 256)
  // EPI_DZ_BWD, top-k mode (D8): kept mask of the forward, [ceil(M / 32)][N] u32, bit (row % 32) of word
  // [row / 32][col] = neuron `row` of sample `col` survived.  Null = off.
  const uint32_t* keep_bits;
  // grouped modes (MoE experts); all null / 0 otherwise
  const int* m_blocks_dev;     // grouped-M: number of valid 128-row blocks (device scalar)
  const int* group_of_mblock;  // grouped-M: expert of each 128-row block
  int b_k_group_stride;        // grouped-M: B k-coordinate offset per expert (weights stacked along k)
  int b_n_group_stride;        // grouped-M: B row (n) coordinate offset per expert (weights stacked along n)
  int bias_group_stride;       // grouped-M: bias_col offset per expert
  const int* seg_offsets;      // grouped-K: [groups + 1] row offsets (multiples of BK) of the experts' segments
  int groups;                  // grouped-K: number of experts
  long long out_group_stride;  // grouped-K: out_f32 offset per expert (elements)
  uint64_t hintA, hintB;
  const float* bias_col;    // [N]  EPI_STORE (may be null)
  const float* bias_row;    // [M]  EPI_UP_FWD (may be null)
  const int8_t* assign;     // [M]  EPI_UP_FWD / EPI_DZ_BWD: 0 relu, 1 tanh, else binary step
  float* partials;          // [PARTIALS_PER_TILE * n_blocks][3][M]  EPI_UP_FWD / EPI_DZ_BWD (may be null)
  float* out_f32;           // EPI_F32
  long long ldc;            // EPI_F32 row pitch in elements
  const float* row_scale;   // EPI_F32 (may be null)
  const float* col_scale;   // EPI_F32 (may be null)
  float scale;              // EPI_F32
  const float* scale_dev;   // EPI_F32: optional device scalar multiplied into scale (may be null)
  int atomic_out;           // EPI_F32: 1 = red.add, 0 = store
  float gate_generalist;    // 1.0
  float gate_specialist;    // 0.1
};

struct WorkItem {
  int m0, n0, nb, kb0, kb1;
  int group;  // expert (grouped modes), else 0
};

// `tile_m` = rows of D per work item (128, or 256 for a CTA pair); `m_off` = this CTA's row offset inside it.
// `m_blocks` is passed in because grouped-M kernels read it from device memory.
__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int t, int tile_m, int m_off, int m_blocks) {
  WorkItem w;
  w.group = 0;
  if (p.seg_offsets != nullptr) {  // grouped-K: t = group * tiles + tile, the whole segment is one reduction
    const int n_tiles = m_blocks * p.n_blocks;
    w.group = t / n_tiles;
    const int tile = t - w.group * n_tiles;
    const int nb = tile % p.n_blocks, mb = tile / p.n_blocks;
    w.m0 = mb * tile_m + m_off;
    w.n0 = nb * BN;
    w.nb = nb;
    w.kb0 = p.seg_offsets[w.group] / BK;
    w.kb1 = p.seg_offsets[w.group + 1] / BK;
    return w;
  }
  const int n_tiles = m_blocks * p.n_blocks;
  const int tile = t % n_tiles;
  const int ks = t / n_tiles;
  int mb, nb;
  if (p.raster_m_fast) {
    mb = tile % m_blocks;
    nb = tile / m_blocks;
  } else {
    nb = tile % p.n_blocks;
    mb = tile / p.n_blocks;
  }
  w.m0 = mb * tile_m + m_off;
  w.n0 = nb * BN;
  w.nb = nb;
  w.kb0 = ks * p.kb_per_split;
  w.kb1 = min(w.kb0 + p.kb_per_split, p.k_blocks);
  if (p.group_of_mblock != nullptr) w.group = p.group_of_mblock[mb];
  return w;
}

// Per-thread activation constants derived from the neuron's assignment (README.md:51-56).
struct ActConst {
  bool is_tanh;
  bool is_step;
  float gate;
};
__device__ __forceinline__ ActConst make_act(int a, float g_gen, float g_spec) {
  ActConst c;
  c.is_tanh = (a == 1);
  c.is_step = (a != 0 && a != 1);  // anything outside {0,1} falls to the last branch of the nested where
  c.gate = c.is_step ? g_spec : g_gen;
  return c;
}

// Warp-level specialisation of the fused epilogues.  The multiplex is per NEURON (= per thread), so a warp whose 32
// neurons share one assignment runs a branch-free body; brain maps are spatially smoothed (smoothing.py:30-91), so
// that is the common case, and the all-generalist map every layer starts with (D7) is entirely WM_RELU.
enum WarpMode : int { WM_RELU = 0, WM_STEP = 1, WM_MIXED = 2 };

struct RunningSums {
  float s_abs, s_sum, s_sq;
};

// forward, 8 accumulator columns -> 4 packed z words + 4 packed h words; MODE is warp-uniform
// SIGNS: also collect (z > 0) into `bits` (interleaved: the pair's two columns land 16 bits apart).
template <int MODE, bool MASKED, bool SIGNS>
__device__ __forceinline__ void up_fwd_8(const uint32_t* v, float bias_r, const ActConst& act, bool warp_has_tanh,
                                         int jcol, int nvalid, uint32_t* zw, uint32_t* hw, RunningSums& rs,
                                         uint32_t& bits) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t zp = pack_bf16x2(__uint_as_float(v[2 * i]) + bias_r, __uint_as_float(v[2 * i + 1]) + bias_r);
    const float z0 = bf16lo_to_f32(zp), z1 = bf16hi_to_f32(zp);
    if (SIGNS) bits |= (bf16x2_gt_zero_mask(zp) & 0x00010001u) << (((jcol & 31) >> 1) + i);
    float h0 = 0.f, h1 = 0.f;
    if (MODE == WM_RELU) {          // h = relu(z) + 1.0 * z: packed, below
    } else if (MODE == WM_STEP) {   // h = step(z) + gate * z
      h0 = fmaf(act.gate, z0, z0 > 0.f ? 1.f : 0.f);
      h1 = fmaf(act.gate, z1, z1 > 0.f ? 1.f : 0.f);
    } else {
      // F_a(z): relu -> z if z>0 ; binary_step -> 1 if z>0 ; else 0.  tanh lanes overwrite.
      float f0 = (z0 > 0.f) ? (act.is_step ? 1.f : z0) : 0.f;
      float f1 = (z1 > 0.f) ? (act.is_step ? 1.f : z1) : 0.f;
      if (warp_has_tanh) {
        const float t0 = tanh_approx(z0), t1 = tanh_approx(z1);
        f0 = act.is_tanh ? t0 : f0;
        f1 = act.is_tanh ? t1 : f1;
      }
      h0 = fmaf(act.gate, z0, f0);
      h1 = fmaf(act.gate, z1, f1);
    }
    zw[i] = zp;
    hw[i] = (MODE == WM_RELU) ? bf16x2_relu_plus_self(zp) : pack_bf16x2(h0, h1);
    if (!MASKED) {
      rs.s_abs += fabsf(z0) + fabsf(z1);
      rs.s_sum += z0 + z1;
      rs.s_sq = fmaf(z0, z0, rs.s_sq);
      rs.s_sq = fmaf(z1, z1, rs.s_sq);
    } else {  // ragged last chunk: samples beyond N are not counted
      if (jcol + 2 * i < nvalid) { rs.s_abs += fabsf(z0); rs.s_sum += z0; rs.s_sq = fmaf(z0, z0, rs.s_sq); }
      if (jcol + 2 * i + 1 < nvalid) { rs.s_abs += fabsf(z1); rs.s_sum += z1; rs.s_sq = fmaf(z1, z1, rs.s_sq); }
    }
  }
}

// backward, 8 columns: dz = acc * (F_a'(z) + gate_a).  Samples beyond N have acc == 0 exactly (TMA zero fill), so
// they add nothing to the sums and need no masking.  m_pos / m_neg are the per-neuron multipliers for z > 0 / z <= 0:
// relu (1 + gate, gate), binary step with the straight-through estimator (1 + gate both), tanh computed per element.
// KEEP: top-k activation sparsification (D8) -- bit c of `keep8` says whether column c of these 8 survived the
// forward's per-token top-k; dropped entries have h = 0 downstream, so their dL/dz is zero.
template <bool HAS_TANH, bool KEEP>
__device__ __forceinline__ void dz_bwd_8(const uint32_t* v, const uint32_t* zw, float m_pos, float m_neg,
                                         const ActConst& act, uint32_t* dw, RunningSums& rs, uint32_t keep8) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float z0 = bf16lo_to_f32(zw[i]), z1 = bf16hi_to_f32(zw[i]);
    float m0 = z0 > 0.f ? m_pos : m_neg;
    float m1 = z1 > 0.f ? m_pos : m_neg;
    if (HAS_TANH) {
      const float t0 = tanh_approx(z0), t1 = tanh_approx(z1);
      m0 = act.is_tanh ? (fmaf(-t0, t0, 1.f) + act.gate) : m0;
      m1 = act.is_tanh ? (fmaf(-t1, t1, 1.f) + act.gate) : m1;
    }
    if (KEEP) {
      m0 = (keep8 >> (2 * i)) & 1u ? m0 : 0.f;
      m1 = (keep8 >> (2 * i + 1)) & 1u ? m1 : 0.f;
    }
    const uint32_t dp = pack_bf16x2(__uint_as_float(v[2 * i]) * m0, __uint_as_float(v[2 * i + 1]) * m1);
    dw[i] = dp;
    const float d0 = bf16lo_to_f32(dp), d1 = bf16hi_to_f32(dp);
    rs.s_abs += fabsf(d0) + fabsf(d1);
    rs.s_sum += d0 + d1;
    rs.s_sq = fmaf(d0, d0, rs.s_sq);
    rs.s_sq = fmaf(d1, d1, rs.s_sq);
  }
}

// backward from the sign bits (no tanh neuron in the warp): bit i of `bits8` = (z > 0) for column i of these 8
// (interleaved layout: bits i and 16 + i for columns 2 i and 2 i + 1)
template <bool KEEP>
__device__ __forceinline__ void dz_bwd_8_bits(const uint32_t* v, uint32_t bits8, float m_pos, float m_neg,
                                              uint32_t* dw, RunningSums& rs, uint32_t keep8) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m0 = (bits8 & (1u << i)) ? m_pos : m_neg;
    float m1 = (bits8 & (1u << (16 + i))) ? m_pos : m_neg;
    if (KEEP) {
      m0 = (keep8 >> (2 * i)) & 1u ? m0 : 0.f;
      m1 = (keep8 >> (2 * i + 1)) & 1u ? m1 : 0.f;
    }
    const uint32_t dp = pack_bf16x2(__uint_as_float(v[2 * i]) * m0, __uint_as_float(v[2 * i + 1]) * m1);
    dw[i] = dp;
    const float d0 = bf16lo_to_f32(dp), d1 = bf16hi_to_f32(dp);
    rs.s_abs += fabsf(d0) + fabsf(d1);
    rs.s_sum += d0 + d1;
    rs.s_sq = fmaf(d0, d0, rs.s_sq);
    rs.s_sq = fmaf(d1, d1, rs.s_sq);
  }
}

// Per-thread view of the work-item sequence of this CTA (pair).  Static: t, t + step, ...  Dynamic: the j-th call of
// advance() consumes the j-th cluster-launch-control response (slot j % CLC_SLOTS, phase (j / CLC_SLOTS) & 1).
struct TileSched {
  int t;            // current work item
  int step, total;  // static schedule
  bool clc;
  uint32_t j;       // responses consumed (= index of the current item in this cluster's sequence)
  uint64_t* full;   // [CLC_SLOTS] local: response j has landed
  uint64_t* empty;  // [CLC_SLOTS] in the LEADER CTA: every consumer of both CTAs has read response j
  uint32_t resp;    // shared address of the response ring
  // the first item: every launched cluster of a dynamic grid owns one; a static worker may have none (grouped-M
  // kernels size their grid for the row capacity and read the number of valid row blocks from device memory)
  __device__ __forceinline__ bool first_valid() const { return clc || t < total; }
  // `arriver`: this thread releases the slot (exactly one thread per consumer agent); warp_collective: every lane
  // of the warp calls this together
  __device__ __forceinline__ bool advance(bool arriver, bool warp_collective) {
    if (!clc) {
      t += step;
      return t < total;
    }
    const uint32_t s = j % CLC_SLOTS;
    mbar_wait(&full[s], (j / CLC_SLOTS) & 1u);
    int x;
    const bool valid = clc_read(resp + 16u * s, x);
    fence_proxy_async_smem();  // our (generic-proxy) read is ordered before the next async-proxy write of this slot
    if (warp_collective) __syncwarp();
    if (arriver) mbar_arrive_leader(&empty[s]);
    ++j;
    t = x >> 1;  // cluster index = first CTA's blockIdx.x / 2
    return valid;
  }
};

template <bool A_MN, bool B_MN, int EPI, bool PAIR>
__global__ void __launch_bounds__(num_threads(EPI, PAIR), 1) forde_gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr int EPI_WARPS = epi_warps(EPI, PAIR);
  constexpr int CPW = (BN / CHUNK) / (EPI_WARPS / 4);  // 64-column chunks per epilogue warp per tile
  constexpr int STAGES = gemm_stages(EPI, PAIR);
  constexpr int WARP_STAGING = epi_warp_staging_bytes(EPI, PAIR);
  constexpr int STAGE_BYTES = stage_bytes(PAIR);
  constexpr int BNL = PAIR ? BN / 2 : BN;       // B rows (D columns) staged by this CTA
  constexpr int TILE_M = PAIR ? 2 * BM : BM;    // D rows per work item

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(staging + EPI_WARPS * WARP_STAGING);  // [BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + BN * 4);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  constexpr int ZR = dz_ring(PAIR), ZD = dz_dist(PAIR);
  uint64_t* z_bar = tempty_bar + 2;          // [EPI_WARPS][ZR]
  uint64_t* clc_full = z_bar + 2 * 16;       // [CLC_SLOTS]
  uint64_t* clc_empty = clc_full + CLC_SLOTS;  // [CLC_SLOTS] (the leader CTA's are used)
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(clc_empty + CLC_SLOTS);
  static_assert(2 * 6 + 4 + 2 * 16 + 2 * CLC_SLOTS <= 62, "barrier block is 512 bytes");
  const uint32_t clc_resp = smem_u32(reinterpret_cast<uint8_t*>(bars) + 512);  // [CLC_SLOTS] x 16 B, 16-byte aligned
  const bool clc = PAIR && p.clc != 0;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.m_blocks_dev != nullptr) ? *p.m_blocks_dev : p.m_blocks;
  const int total_work = (p.seg_offsets != nullptr) ? p.groups * m_blocks * p.n_blocks : m_blocks * p.n_blocks * p.split_k;
  const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_off = cta_rank * BM;
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // every role walks the same item sequence through its own copy of this
  const TileSched sched0{worker, num_workers, total_work, clc, 0u, clc_full, clc_empty, clc_resp};

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (EPI != EPI_F32) tma_prefetch_desc(&p.tmC0);
    if (epi_two_out(EPI)) tma_prefetch_desc(&p.tmC1);
    if (epi_z_ring(EPI)) tma_prefetch_desc(&p.tmZ);
    if (EPI == EPI_DZ_BWD) tma_prefetch_desc(&p.tmH);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // pair: both CTAs' epilogues release the leader
    }
    for (int i = 0; i < ZR * EPI_WARPS; ++i) mbar_init(&z_bar[i], 1);
    for (int i = 0; i < CLC_SLOTS; ++i) {
      mbar_init(&clc_full[i], 1);
      mbar_init(&clc_empty[i], 2 + 1 + 2 * EPI_WARPS);  // both producers, the MMA issuer, both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_pair(tmem_ptr_s, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_s, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      TileSched sc = sched0;
      for (bool more = sc.first_valid(); more; more = sc.advance(true, false)) {
        const int t = sc.t;
        if (clc) {
          // ask for the NEXT item now, one item ahead: the answer is back long before this item's loads are issued
          const uint32_t s = sc.j % CLC_SLOTS;
          if (cta_rank == 0) mbar_wait(&clc_empty[s], ((sc.j / CLC_SLOTS) & 1u) ^ 1u);  // slot read by everyone
          mbar_expect_tx(&clc_full[s], 16);  // each CTA arms its own barrier; the response is multicast to both
          if (cta_rank == 0) clc_try_cancel_multicast(clc_resp + 16u * s, &clc_full[s]);
        }
        const WorkItem w = decode_work(p, t, TILE_M, m_off, m_blocks);
        // first B row (D column) this CTA stages (+ the expert's slice of n-stacked weights)
        const int nb0 = w.n0 + (PAIR ? cta_rank * BNL : 0) + w.group * p.b_n_group_stride;
        const int kgo = w.group * p.b_k_group_stride;  // expert's slice of k-stacked weights
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // pair: both CTAs' bytes are counted on the leader's barrier; only the leader arms it
          if (!PAIR || cta_rank == 0) mbar_expect_tx(&full_bar[stage], (PAIR ? 2 : 1) * STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          const int k0 = kb * BK;
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1, uint64_t hint) {
            if (PAIR) tma_load_2d_pair(tm, &full_bar[stage], dst, c0, c1, hint);
            else tma_load_2d(tm, &full_bar[stage], dst, c0, c1, hint);
          };
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) load(&p.tmA, sa + j * (BK * 128), w.m0 + j * 64, k0, p.hintA);
          } else {
            load(&p.tmA, sa, k0, w.m0, p.hintA);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) load(&p.tmB, sb + j * (BK * 128), nb0 + j * 64, k0 + kgo, p.hintB);
          } else {
            load(&p.tmB, sb, k0 + kgo, nb0, p.hintB);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN, A_MN, B_MN);
    // K-major tile: rows of 128 B, 8-row swizzle atoms 1024 B apart; one UMMA_K step = 32 B inside the row.
    // MN-major tile: 64-element (128 B) row per k, 8-k atoms 1024 B apart (SBO), 64-wide MN chunks
    //                BK*128 B apart (LBO); one UMMA_K step = 16 k-rows = 2048 B.
    constexpr uint32_t a_lbo = A_MN ? BK * 128 : 0, b_lbo = B_MN ? BK * 128 : 0;
    constexpr uint32_t a_kstep = A_MN ? UMMA_K * 128 : UMMA_K * 2, b_kstep = B_MN ? UMMA_K * 128 : UMMA_K * 2;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (!PAIR || cta_rank == 0) {  // in a pair only the leader issues; its MMAs drive both SMs' tensor cores
      TileSched sc = sched0;
      for (bool more = sc.first_valid(); more; more = sc.advance(lane == 0, true), ++it) {
        const int t = sc.t;
        const WorkItem w = decode_work(p, t, TILE_M, m_off, m_blocks);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        if (w.kb0 >= w.kb1 && elect_one()) {  // empty reduction: nothing to wait for, the epilogue writes zeros
          if (PAIR) umma_commit_pair(&tfull_bar[as]);
          else umma_commit(&tfull_bar[as]);
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
            const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(a_base + k * a_kstep, a_lbo, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(b_base + k * b_kstep, b_lbo, 1024);
              if (PAIR) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[stage]);  // frees the stage in BOTH CTAs
              if (kb == w.kb1 - 1) umma_commit_pair(&tfull_bar[as]);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == w.kb1 - 1) umma_commit(&tfull_bar[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;          // TMEM lane quarter owned by this warp
    const int ew = warp - 2;         // staging region index
    const int e = ew >> 2;           // which column half of the tile this warp drains
    uint8_t* wst = staging + ew * WARP_STAGING;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    int it = 0;

    // z^T prefetch state (EPI_DZ_BWD): g = running chunk index of this warp (tile g / CPW, chunk g % CPW);
    // chunk g lives in box g % ZR and dz is written over it in place.
    int g = 0;
    uint32_t zph = 0;  // phase bit per box: a box's barrier advances only when a load into it has been consumed
    uint64_t* zb = z_bar + ew * ZR;
    // Warp-collective.  The backward only needs z itself for tanh neurons (F' = 1 - tanh^2 z); for relu neurons the
    // sign of h = z + relu(z) is the sign of z and the straight-through step needs nothing, so a warp without tanh
    // neurons reads the h^T box (which the forward stores anyway) and the forward never has to write its z^T box.
    // item sequence of this warp, read one item ahead (the z^T prefetch of a tile's last chunk targets the next item)
    TileSched sc = sched0;
    int cur_t = sc.first_valid() ? sc.t : -1;
    int next_t = (cur_t >= 0 && sc.advance(lane == 0, true)) ? sc.t : -1;
    auto z_prefetch = [&](int gi, bool wait_for_store) {
      const int nt = (gi / CPW == it) ? cur_t : next_t;  // chunk gi lies in the current item or in the next one
      if (nt < 0) return;
      const WorkItem wn = decode_work(p, nt, TILE_M, m_off, m_blocks);
      bool use_z = true;
      if (EPI == EPI_DZ_BWD) {
        const int rn = wn.m0 + q * 32 + lane;
        use_z = __any_sync(0xffffffffu, rn < p.M && p.assign[rn] == 1);
        if (!use_z && p.sign_bits != nullptr) return;  // sign-bit path: this chunk loads nothing into its box
      }
      if (lane == 0) {
        // the dz store that last read this box (chunk gi - ZR) must have drained
        if (wait_for_store) tma_wait_read<ZR - ZD - 1>();
        const int bi = gi % ZR;
        mbar_expect_tx(&zb[bi], BOX_BYTES);
        tma_load_2d(use_z ? &p.tmZ : &p.tmH, &zb[bi], wst + bi * BOX_BYTES, wn.n0 + (e * CPW + gi % CPW) * CHUNK,
                    wn.m0 + q * 32, kEvictFirst);
      }
    };
    if (epi_z_ring(EPI)) {
#pragma unroll
      for (int i = 0; i < ZD; ++i) z_prefetch(i, false);
    }

    for (; cur_t >= 0; ++it) {
      const int t = cur_t;
      const WorkItem w = decode_work(p, t, TILE_M, m_off, m_blocks);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row = w.m0 + q * 32 + lane;

      // per-tile, per-thread constants
      float bias_r = 0.f;
      ActConst act = make_act(0, p.gate_generalist, p.gate_specialist);
      bool warp_has_tanh = false;
      int wmode = WM_MIXED;
      bool store_z = true;
      float m_pos = 0.f, m_neg = 0.f;
      if (epi_neuron_rows(EPI)) {
        int a = 0;
        if (row < p.M) {
          a = p.assign[row];
          if (EPI == EPI_UP_FWD && p.bias_row != nullptr) bias_r = p.bias_row[row];
        }
        act = make_act(a, p.gate_generalist, p.gate_specialist);
        warp_has_tanh = __any_sync(0xffffffffu, act.is_tanh);
        const bool relu_row = !act.is_tanh && !act.is_step;
        wmode = __all_sync(0xffffffffu, relu_row) ? WM_RELU : __all_sync(0xffffffffu, act.is_step) ? WM_STEP : WM_MIXED;
        store_z = (p.save_z != 0) || warp_has_tanh;  // see z_prefetch: only tanh neurons need z in the backward
        m_pos = 1.f + act.gate;                      // relu'(z>0) + gate ; STE slope 1 + gate
        m_neg = act.is_step ? 1.f + act.gate : act.gate;
      }
      // sign-bit paths (warp-uniform per tile): the forward writes bits instead of the z box, the backward reads them
      const bool want_bits = (EPI == EPI_UP_FWD) && !store_z && p.sign_bits != nullptr;
      const bool bits_path = (EPI == EPI_DZ_BWD) && !warp_has_tanh && p.sign_bits != nullptr;
      uint32_t mw[4] = {0u, 0u, 0u, 0u};  // this warp's sign words of the tile: [chunk][half]
      static_assert(EPI != EPI_DZ_BWD || CPW <= 2, "sign words per tile");
      if (bits_path && row < p.M) {
        const int w0 = (w.n0 + e * CPW * CHUNK) / 32;  // first sign word of this warp's columns
#pragma unroll
        for (int i = 0; i < 2 * CPW; ++i) mw[i] = __ldg(p.sign_bits + sign_index(row, w0 + i, p.M));
      }
      if (epi_col_bias(EPI)) {
        // stage this tile's column bias (of this tile's expert in the grouped mode) once for the epilogue warps
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // previous tile's bias no longer read
        const int c = threadIdx.x - 64;
        if (c < BN) {  // 8 epilogue warps = 256 threads = BN
          const int col = w.n0 + c;
          bias_s[c] = (p.bias_col != nullptr && col < p.N) ? p.bias_col[w.group * p.bias_group_stride + col] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = lane_taddr + as * BN;

      static_assert(!epi_neuron_rows(EPI) || CPW % CHUNKS_PER_SLOT == 0, "a partials slot must belong to one warp");
      RunningSums rs{0.f, 0.f, 0.f};  // per-neuron sums of the current partials slot (fused epilogues)
      uint32_t va_next[32];           // tmem_pipe == 2: first half of the next chunk, in flight
#pragma unroll 1
      for (int cc = 0; cc < CPW; ++cc) {
        const int c = e * CPW + cc;
        const int col0 = w.n0 + c * CHUNK;
        const int nvalid = p.N - col0;  // valid columns in this chunk (may be <= 0 or >= CHUNK)
        const bool full = nvalid >= CHUNK;

        if (epi_z_ring(EPI)) z_prefetch(g + ZD, true);  // refill the box chunk g + ZD will use
        // single-buffered staging: the previous chunk's TMA stores must have read the boxes (the other epilogue
        // warps of this SM sub-partition cover the wait)
        if ((epi_two_out(EPI) || EPI == EPI_STORE) && lane == 0) tma_wait_read<0>();
        constexpr uint32_t hoff = BOX_BYTES;
        if (epi_z_ring(EPI)) {
          if (bits_path) {  // no load to wait for; the store that last read this box (chunk g - ZR) must have drained
            if (lane == 0) tma_wait_read<ZR - 1>();
          } else {
            mbar_wait(&zb[g % ZR], (zph >> (g % ZR)) & 1u);
            zph ^= 1u << (g % ZR);
          }
        }
        __syncwarp();
        uint8_t* obuf = epi_z_ring(EPI) ? (wst + (g % ZR) * BOX_BYTES) : wst;
        const uint32_t orow_s = smem_u32(obuf) + lane * 128;
        // top-k mode: the kept words of this chunk's 64 samples for this warp's 32-neuron group (lane l holds the
        // words of samples col0 + l and col0 + 32 + l; bit (row % 32) is this thread's neuron)
        uint32_t kw[2] = {0u, 0u};
        const bool use_keep = (EPI == EPI_DZ_BWD) && p.keep_bits != nullptr;
        if (use_keep && w.m0 + q * 32 < p.M) {
          const uint32_t* kp = p.keep_bits + static_cast<long long>((w.m0 + q * 32) >> 5) * p.N;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int col = col0 + hh * 32 + lane;
            kw[hh] = col < p.N ? __ldg(kp + col) : 0u;
          }
        }

        // the 64-column chunk is drained in two 32-column halves to keep the live registers low
        auto drain_half = [&](const uint32_t* v_in, const int half) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v_in[j];
          const int jbase = half * 32;  // column offset inside the chunk
          if (EPI == EPI_F32 && w.kb0 >= w.kb1) {  // empty reduction (grouped-K expert without tokens): zeros
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }

          if (EPI == EPI_F32) {
            if (row < p.M && nvalid > jbase) {
              const float rs = (p.row_scale != nullptr ? p.row_scale[row] : 1.f) * p.scale *
                               (p.scale_dev != nullptr ? *p.scale_dev : 1.f);
              float* orow = p.out_f32 + w.group * p.out_group_stride + static_cast<long long>(row) * p.ldc + col0 + jbase;
              const bool vec_ok = (nvalid >= jbase + 32) && ((p.ldc & 3) == 0) &&
                                  ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
              if (vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float f0 = __uint_as_float(v[j]) * rs, f1 = __uint_as_float(v[j + 1]) * rs;
                  float f2 = __uint_as_float(v[j + 2]) * rs, f3 = __uint_as_float(v[j + 3]) * rs;
                  if (p.col_scale != nullptr) {
                    const float4 cs = *reinterpret_cast<const float4*>(p.col_scale + col0 + jbase + j);
                    f0 *= cs.x; f1 *= cs.y; f2 *= cs.z; f3 *= cs.w;
                  }
                  if (p.atomic_out) {
                    red_add_v4_f32(orow + j, f0, f1, f2, f3);
                  } else {
                    *reinterpret_cast<float4*>(orow + j) = make_float4(f0, f1, f2, f3);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (jbase + j < nvalid) {
                    float f = __uint_as_float(v[j]) * rs;
                    if (p.col_scale != nullptr) f *= p.col_scale[col0 + jbase + j];
                    if (p.atomic_out) atomicAdd(orow + j, f);
                    else orow[j] = f;
                  }
                }
              }
            }
          } else if (EPI == EPI_STORE) {
            const float* bsrc = bias_s + c * CHUNK + jbase;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const float4 b0 = *reinterpret_cast<const float4*>(bsrc + j8 * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bsrc + j8 * 8 + 4);
              const uint32_t w0 = pack_bf16x2(__uint_as_float(v[j8 * 8 + 0]) + b0.x, __uint_as_float(v[j8 * 8 + 1]) + b0.y);
              const uint32_t w1 = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]) + b0.z, __uint_as_float(v[j8 * 8 + 3]) + b0.w);
              const uint32_t w2 = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]) + b1.x, __uint_as_float(v[j8 * 8 + 5]) + b1.y);
              const uint32_t w3 = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]) + b1.z, __uint_as_float(v[j8 * 8 + 7]) + b1.w);
              st_shared_v4(orow_s + ((static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4), w0, w1, w2, w3);
            }
          } else if (EPI == EPI_UP_FWD) {
            // z box at wst, h box at wst + BOX_BYTES
            uint32_t sbits = 0u;  // (z > 0) of this half's 32 samples
            auto run = [&](auto mode, auto masked, auto signs) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t zw[4], hw[4];
                up_fwd_8<decltype(mode)::value, decltype(masked)::value, decltype(signs)::value>(
                    v + j8 * 8, bias_r, act, warp_has_tanh, jbase + j8 * 8, nvalid, zw, hw, rs, sbits);
                const uint32_t off = (static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4;
                if (store_z) st_shared_v4(orow_s + off, zw[0], zw[1], zw[2], zw[3]);
                st_shared_v4(orow_s + hoff + off, hw[0], hw[1], hw[2], hw[3]);
              }
            };
            using T = std::true_type;
            using F = std::false_type;
            auto dispatch = [&](auto signs) {
              if (!full) run(std::integral_constant<int, WM_MIXED>{}, T{}, signs);
              else if (wmode == WM_RELU) run(std::integral_constant<int, WM_RELU>{}, F{}, signs);
              else if (wmode == WM_STEP) run(std::integral_constant<int, WM_STEP>{}, F{}, signs);
              else run(std::integral_constant<int, WM_MIXED>{}, F{}, signs);
            };
            if (want_bits) {
              dispatch(T{});
              if (row < p.M) p.sign_bits[sign_index(row, col0 / 32 + half, p.M)] = sbits;
            } else {
              dispatch(F{});
            }
          } else if (EPI == EPI_STORE_ADD) {
            const float* bsrc = bias_s + c * CHUNK + jbase;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint32_t off = (static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4;
              uint32_t cw[4], ow[4];
              ld_shared_v4(orow_s + off, cw[0], cw[1], cw[2], cw[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int j = j8 * 8 + i * 2;
                ow[i] = pack_bf16x2(__uint_as_float(v[j]) + bsrc[j] + bf16lo_to_f32(cw[i]),
                                    __uint_as_float(v[j + 1]) + bsrc[j + 1] + bf16hi_to_f32(cw[i]));
              }
              st_shared_v4(orow_s + off, ow[0], ow[1], ow[2], ow[3]);
            }
          } else if (EPI == EPI_MOE_UP) {
            // z box at wst (tmC0), h = gelu_tanh(z) box at wst + BOX_BYTES (tmC1); flax nn.gelu default (moe.py:27)
            const float* bsrc = bias_s + c * CHUNK + jbase;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint32_t zw[4], hw[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int j = j8 * 8 + i * 2;
                const uint32_t zp = pack_bf16x2(__uint_as_float(v[j]) + bsrc[j], __uint_as_float(v[j + 1]) + bsrc[j + 1]);
                const float z0 = bf16lo_to_f32(zp), z1 = bf16hi_to_f32(zp);
                zw[i] = zp;
                hw[i] = pack_bf16x2(gelu_tanh_f(z0), gelu_tanh_f(z1));
              }
              const uint32_t off = (static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4;
              st_shared_v4(orow_s + off, zw[0], zw[1], zw[2], zw[3]);
              st_shared_v4(orow_s + BOX_BYTES + off, hw[0], hw[1], hw[2], hw[3]);
            }
          } else if (EPI == EPI_MOE_DZ) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint32_t off = (static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4;
              uint32_t zw[4], dw[4];
              ld_shared_v4(orow_s + off, zw[0], zw[1], zw[2], zw[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int j = j8 * 8 + i * 2;
                const float z0 = bf16lo_to_f32(zw[i]), z1 = bf16hi_to_f32(zw[i]);
                dw[i] = pack_bf16x2(__uint_as_float(v[j]) * gelu_tanh_grad_f(z0),
                                    __uint_as_float(v[j + 1]) * gelu_tanh_grad_f(z1));
              }
              st_shared_v4(orow_s + off, dw[0], dw[1], dw[2], dw[3]);
            }
          } else {  // EPI_DZ_BWD: dz over z, in place (same thread, same bytes)
            // kept bits of columns j8 * 8 .. + 7 for this thread's neuron (warp-collective: every lane asks lane j)
            auto keep8_of = [&](int j8) {
              uint32_t k8 = 0u;
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8)
                k8 |= ((__shfl_sync(0xffffffffu, kw[half], j8 * 8 + c8) >> (row & 31)) & 1u) << c8;
              return k8;
            };
            auto run = [&](auto has_tanh, auto keep) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                const uint32_t off = (static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4;
                uint32_t zw[4], dw[4];
                const uint32_t k8 = decltype(keep)::value ? keep8_of(j8) : 0u;
                ld_shared_v4(orow_s + off, zw[0], zw[1], zw[2], zw[3]);
                dz_bwd_8<decltype(has_tanh)::value, decltype(keep)::value>(v + j8 * 8, zw, m_pos, m_neg, act, dw, rs, k8);
                st_shared_v4(orow_s + off, dw[0], dw[1], dw[2], dw[3]);
              }
            };
            auto run_bits = [&](auto keep) {
              uint32_t word = mw[half];
              if (CPW > 1 && cc == 1) word = mw[2 + half];
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t dw[4];
                const uint32_t k8 = decltype(keep)::value ? keep8_of(j8) : 0u;
                dz_bwd_8_bits<decltype(keep)::value>(v + j8 * 8, word >> (j8 * 4), m_pos, m_neg, dw, rs, k8);
                st_shared_v4(orow_s + ((static_cast<uint32_t>(half * 4 + j8) ^ sw) << 4), dw[0], dw[1], dw[2], dw[3]);
              }
            };
            using T = std::true_type;
            using F = std::false_type;
            if (use_keep) {  // warp-uniform
              if (bits_path) run_bits(T{});
              else if (warp_has_tanh) run(T{}, T{});
              else run(F{}, T{});
            } else {
              if (bits_path) run_bits(F{});
              else if (warp_has_tanh) run(T{}, F{});
              else run(F{}, F{});
            }
          }
        };
        if (tmem_pipe(EPI) == 2) {
          // as below, and the next chunk's first half is requested before this chunk's second half is processed
          uint32_t vb[32];
          if (cc == 0) tmem_ld32(taddr + c * CHUNK, va_next);
          tmem_ld32(taddr + c * CHUNK + 32, vb);
          tmem_wait_ld();  // (a no-op in SASS: consumers wait on the LDTM scoreboards, so vb stays in flight)
          drain_half(va_next, 0);
          if (cc + 1 < CPW) tmem_ld32(taddr + (c + 1) * CHUNK, va_next);
          tmem_wait_ld();
          drain_half(vb, 1);
        } else if (tmem_pipe(EPI) == 1) {
          // fused epilogues: the second half's tcgen05.ld is in flight while the first half is processed
          uint32_t va[32], vb[32];
          tmem_ld32(taddr + c * CHUNK, va);
          tmem_wait_ld();
          tmem_ld32(taddr + c * CHUNK + 32, vb);
          drain_half(va, 0);
          tmem_wait_ld();
          drain_half(vb, 1);
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t vh[32];
            tmem_ld32(taddr + c * CHUNK + half * 32, vh);
            tmem_wait_ld();
            drain_half(vh, half);
          }
        }

        if (EPI != EPI_F32) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (EPI != EPI_UP_FWD || store_z) tma_store_2d(&p.tmC0, obuf, col0, w.m0 + q * 32);
            if (epi_two_out(EPI)) tma_store_2d(&p.tmC1, obuf + (EPI == EPI_UP_FWD ? hoff : BOX_BYTES), col0, w.m0 + q * 32);
            tma_commit();
          }
        }
        if (epi_z_ring(EPI)) ++g;

        if (epi_neuron_rows(EPI) && ((c + 1) % CHUNKS_PER_SLOT) == 0) {  // slot complete
          if (p.partials != nullptr && row < p.M) {
            float* dst = p.partials + (static_cast<long long>(w.nb * PARTIALS_PER_TILE + c / CHUNKS_PER_SLOT) * 3) * p.M + row;
            dst[0] = rs.s_abs;
            dst[p.M] = rs.s_sum;
            dst[2 * static_cast<long long>(p.M)] = rs.s_sq;
          }
          rs = RunningSums{0.f, 0.f, 0.f};
        }
      }

      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
      cur_t = next_t;
      if (cur_t >= 0) next_t = sc.advance(lane == 0, true) ? sc.t : -1;
    }
    if (EPI != EPI_F32 && lane == 0) tma_wait_all<0>();
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // neither CTA may leave while the other can still touch its smem / TMEM / barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace forde
