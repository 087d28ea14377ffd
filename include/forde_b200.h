/* forde_b200.h -- C ABI of the B200-native FORDE fast-loop hot path.
 *
 * This is the boundary a jax.ffi / XLA custom-call shim (or ctypes, as in this repository) binds.
 * The reference has no native layer; every entry point names the reference code (or README spec)
 * it stands in for.  Conventions for every function:
 *   - `stream` is a cudaStream_t passed as void*; kernels are only ENQUEUED on it.  No entry point
 *     synchronises, allocates device memory or touches host memory other than its arguments.
 *   - the caller owns and pre-allocates every buffer (inputs, outputs, scratch).
 *   - all matrices are row-major; bf16 matrices need 16-byte aligned base pointers and a row pitch
 *     (in elements) that is a multiple of 8.
 *   - return value: 0 on success, a negative FORDE_E* code on a shape / alignment / launch error.
 *     An XLA FFI handler maps non-zero to XLA_FFI_Error.
 *
 * Data layout of the StatefulLayer (reference spec: /root/reference/README.md:47-56):
 *   x, y, dy, dx : [M][d]  bf16, one row per token ("sample")
 *   w_up         : [d][H]  bf16  (Flax Dense `kernel`, (in, out))
 *   w_down       : [H][d]  bf16
 *   b_up [H], b_down [d] : fp32
 *   assign       : [H] int8, 0 = relu (generalist), 1 = tanh (pooling), anything else = binary step
 *   zT, hT, dzT  : [H][ldz] bf16, NEURON-major (one row per neuron, samples contiguous), ldz >= M.
 *                  Keeping the per-neuron axis on the TMEM lanes is what lets the sensing statistics be
 *                  accumulated per thread with no cross-lane reduction.
 *   partials     : [P][3][H] fp32 with P = forde_sl_num_partials(M): per 128-sample block the per-neuron
 *                  sum|v|, sum v, sum v^2 (v = z in the forward, v = dL/dz in the backward).  Written,
 *                  not accumulated: every slot has exactly one writer, so results are deterministic.
 */
#ifndef FORDE_B200_H_
#define FORDE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FORDE_OK 0
#define FORDE_EINVAL (-1)   /* null pointer / non-positive dimension */
#define FORDE_EALIGN (-2)   /* pointer or pitch alignment violated */
#define FORDE_ECUDA (-3)    /* CUDA runtime / driver error (launch, tensor-map encode) */
#define FORDE_ENODEV (-4)   /* no sm_100 device */

int forde_version(void);
const char* forde_error_string(int code);
/* last CUDA error text recorded by this library on the calling thread ("" if none) */
const char* forde_last_cuda_error(void);
int forde_sm_count(void);
/* kernels launched by this library in this process so far (bench.py reports the per-step difference) */
unsigned long long forde_launch_count(void);

/* number of 128-sample partial blocks the forward / backward kernels write for M samples
 * (2 per 256-sample tile column; the last one may be empty) */
int forde_sl_num_partials(int M);

/* ---- generic GEMM  D[M][N] = sum_k A(m,k) * B(n,k)  (test + building block) -------------------------
 * a_mn_major = 0: A stored [M][K] (pitch lda);  1: A stored [K][M] (pitch lda).  Same for B with N. */
int forde_gemm_bf16(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                    long long ldb, void* D, long long ldd, const float* bias_col, int M, int N, int K);
/* D = A . B^T + bias_col + Cin, bf16; Cin has D's shape and pitch and may be D itself (accumulate in place) */
int forde_gemm_bf16_add(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                        long long ldb, void* D, long long ldd, const void* Cin, const float* bias_col, int M, int N,
                        int K);
int forde_gemm_f32(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                   long long ldb, float* D, long long ldd, const float* row_scale, const float* col_scale,
                   float scale, int accumulate, int M, int N, int K);

/* ---- StatefulLayer forward (README.md:47-56; sensing.py:57-59 for the activation moments) ------------
 * up:   zT = (x . w_up + b_up)^T ; hT = F_assign(zT) + gate_assign * zT ; act_partials (may be NULL).
 *       save_z = 1 writes all of zT.  save_z = 0 writes only the 32-neuron x 64-sample boxes whose neuron group
 *       contains a tanh neuron: the backward needs z itself only there (relu: sign(h) == sign(z); binary step:
 *       straight-through), reads hT everywhere else, and the forward saves up to half of its HBM writes.  The
 *       activation statistics are taken from z in registers either way.
 * down: y = h . w_down + b_down */
int forde_sl_up_fwd(void* stream, const void* x, const void* w_up, const float* b_up, const int8_t* assign,
                    void* zT, void* hT, float* act_partials, int M, int d, int H, long long ldz, int save_z);
int forde_sl_down_fwd(void* stream, const void* hT, const void* w_down, const float* b_down, void* y, int M,
                      int d, int H, long long ldz);
/* y = h . w_down + b_down + residual (residual [M][d] bf16, may be y itself): the FORDE block's `x + layer(x)`
 * (README.md:43) formed in the GEMM epilogue, so the add/norm that follows streams one tensor instead of two */
int forde_sl_down_fwd_add(void* stream, const void* hT, const void* w_down, const float* b_down, const void* residual,
                          void* y, int M, int d, int H, long long ldz);
/* Sign-bit variant of the save_z = 0 forward (what the module uses by default).  Where the z box is not written the
 * kernel also records (z > 0), one bit per neuron and sample, in sign_bits: H * forde_sl_sign_words(M) u32 words,
 * 16-byte aligned, word-major [forde_sl_sign_words(M)][H] so that a warp (32 consecutive neurons) touches one
 * 128-byte line per access.  forde_sl_dz_bwd_signs then reads these 1/16 as many bytes instead of the hT box.  Same
 * results as the hT path bit for bit (relu'(z) needs only the sign of z; the straight-through step needs nothing).
 * Measured at C2 vision dims: dz_bwd 191 -> 163 us at no cost to the forward (profiles/r2_ab1_sign_bits.txt). */
int forde_sl_sign_words(int M);
/* Layout of sign_bits, a bit mask (this build returns 3).  Bit 0: bit order inside a 32-sample word, 1 = interleaved
 * (sample 2k -> bit k, sample 2k + 1 -> bit 16 + k: the packed bf16x2 compare yields a pair's bits 16 apart).
 * Bit 1: 1 = word-major [words][H].  The buffer is a private residual between the two kernels; only tools that
 * inspect it need this. */
int forde_sl_sign_layout(void);
int forde_sl_up_fwd_signs(void* stream, const void* x, const void* w_up, const float* b_up, const int8_t* assign,
                          void* zT, void* hT, uint32_t* sign_bits, float* act_partials, int M, int d, int H,
                          long long ldz);
/* both of the above */
int forde_sl_fwd(void* stream, const void* x, const void* w_up, const float* b_up, const void* w_down,
                 const float* b_down, const int8_t* assign, void* zT, void* hT, void* y, float* act_partials,
                 int M, int d, int H, long long ldz, int save_z);

/* ---- StatefulLayer backward (autodiff of the above with the straight-through binary step, README.md:55;
 *      gradient statistics are those of dL/dz, tests/reproduce_grad_sink.py:16-17, sensing.py:62-63) -----
 * dz:      dzT = (dy . w_down^T)^T * (F'(zT) + gate) ; grad_partials (sum|dz|, sum dz (= db_up), sum dz^2);
 *          zT is read only where forde_sl_up_fwd(save_z = 0) would have written it, hT elsewhere; `assign` must be
 *          the assignment the forward ran with
 * dx:      dx = dz . w_up^T
 * dw_up:   dw_up[d][H]   += x^T . dz      (fp32, accumulated with red.add: caller zeroes or carries grads)
 * dw_down: dw_down[H][d] += h^T . dy
 * db_down: db_down[d]    += sum_m dy */
int forde_sl_dz_bwd(void* stream, const void* dy, const void* w_down, const void* zT, const void* hT,
                    const int8_t* assign, void* dzT, float* grad_partials, int M, int d, int H, long long ldz);
/* backward of forde_sl_up_fwd_signs: zT where a neuron group has a tanh neuron, sign_bits elsewhere */
int forde_sl_dz_bwd_signs(void* stream, const void* dy, const void* w_down, const void* zT,
                          const uint32_t* sign_bits, const int8_t* assign, void* dzT, float* grad_partials, int M,
                          int d, int H, long long ldz);
int forde_sl_dx_bwd(void* stream, const void* dzT, const void* w_up, void* dx, int M, int d, int H,
                    long long ldz);
/* dx = dz . w_up^T + residual ([M][d] bf16; may be dx itself): when the layer sits in `x + StatefulLayer(x)` (README.md:43)
 * the skip path's gradient is added in the GEMM epilogue instead of by a separate pass */
int forde_sl_dx_bwd_add(void* stream, const void* dzT, const void* w_up, const void* residual, void* dx, int M, int d,
                        int H, long long ldz);
int forde_sl_dw_up(void* stream, const void* x, const void* dzT, float* dw_up, int M, int d, int H,
                   long long ldz);
int forde_sl_dw_down(void* stream, const void* hT, const void* dy, float* dw_down, int M, int d, int H,
                     long long ldz);
int forde_colsum_bf16(void* stream, const void* a, long long lda, float* out, int rows, int cols);
/* all of the above; db_up[H] += sum dz is taken from grad_partials */
int forde_sl_bwd(void* stream, const void* x, const void* dy, const void* w_up, const void* w_down,
                 const void* zT, const void* hT, const int8_t* assign, void* dzT, void* dx, float* dw_up,
                 float* db_up, float* dw_down, float* db_down, float* grad_partials, int M, int d, int H,
                 long long ldz);

/* ---- optional top-k activation sparsification (nearest reference code: lax.top_k at moe.py:130) ------
 * sparsify: per token (column of hT) keep the k largest |h| over the H neurons and zero the rest IN PLACE (call it
 *   between forde_sl_up_fwd and forde_sl_down_fwd; one read and one write of hT).  idx_out [M][k] i32 = kept neuron
 *   indices in jax.lax.top_k order of |h| (descending, equal values -> lower index first; may be NULL); hoyer_out [M]
 *   f32 = Hoyer sparsity of the dense h of each token (may be NULL); keep_bits [ceil(H/32)][M] u32 = kept mask.
 * dz_bwd_topk: forde_sl_dz_bwd / forde_sl_dz_bwd_signs with that mask applied in the same epilogue: dL/dz = 0 where the
 *   forward dropped (h, m), and grad_partials are the sums of the MASKED dL/dz -- no extra pass over dzT.  Exactly one
 *   of hT / sign_bits is the source of sign(z) for the neuron groups without a tanh neuron (the other may be NULL;
 *   sign_bits wins when both are given). */
int forde_sl_topk_sparsify(void* stream, void* hT, long long ldz, int M, int H, int k, int* idx_out, float* hoyer_out,
                           uint32_t* keep_bits);
int forde_sl_dz_bwd_topk(void* stream, const void* dy, const void* w_down, const void* zT, const void* hT,
                         const uint32_t* sign_bits, const uint32_t* keep_bits, const int8_t* assign, void* dzT,
                         float* grad_partials, int M, int d, int H, long long ldz);

/* ---- sensing finalisation (sensing.py:5-69) --------------------------------------------------------
 * Reduces the partials in a fixed order and produces this step's (H,5) statistics
 *   [grad_gini, grad_gdp, act_gini, act_gdp, act_variance]   (gini = Hoyer sparsity over the M samples)
 * then accumulates the stats_buffer leaves (moe.py:265-279 pattern): neuron_stats (H,5) += step stats,
 * moments (H,5) += [sum|z|, sum z, sum z^2, sum|g|, sum g^2], sample_count += M, step_count += 1.
 * moments_lo (H,5; may be NULL) makes the accumulator a double-single pair: moments holds the fp32 high part,
 * moments_lo the rounding residue, so sums over a whole slow-loop interval keep ~48 bits (the pooled variance
 * sum z^2 / n - mean^2 cancels catastrophically in plain fp32 when |mean| >> std).  Both remain fp32 leaves.
 * step_stats / neuron_stats / moments / counters may each be NULL.  db_up (may be NULL) += sum g. */
int forde_stats_finalize(void* stream, const float* act_partials, const float* grad_partials, int num_partials,
                         int H, int M, float* step_stats, float* neuron_stats, float* moments, float* moments_lo,
                         int* sample_count, int* step_count, float* db_up);
/* stats from already pooled moments, given in fp64 = double(moments) + double(moments_lo), summed over ranks
 * (used after the cross-rank all-reduce of the accumulators) */
int forde_stats_from_moments(void* stream, const double* moments, const int* sample_count, int H,
                             float* stats_out);

/* ---- CLIP contrastive loss (README.md:35) ------------------------------------------------
 * img [Bl][E], txt [Bg][E] bf16 (txt already all-gathered; Bl = Bg on one GPU), row_offset = global index
 * of local image row 0.  log_temp is a DEVICE scalar (learnable log temperature tau), so no host sync.
 * fwd:  u = img/|img|, v = txt/|txt| (bf16), logit_scale_out[1] = exp(tau),
 *       S = exp(tau) * u . v^T (fp32 [Bl][Bg]), row_lse[Bl] (complete: a rank owns whole rows),
 *       col_max[Bg], col_sum[Bg] (max / sum-exp over the LOCAL rows), diag[Bl] = S[i][row_offset+i].
 * Across ranks: all-reduce(MAX) col_max -> g; forde_clip_col_rescale(col_max_local, g, col_sum);
 *       all-reduce(SUM) col_sum.  On one GPU skip both.
 * loss: col_lse = col_max + log col_sum;  loss[1] += this rank's share of
 *       1/(2 Bg) (sum_i (row_lse_i - diag_i) + sum_j (col_lse_j - diag_j))   (all-reduce(SUM) across ranks) */
int forde_clip_fwd(void* stream, const void* img, const void* txt, const float* log_temp, int Bl, int Bg, int E,
                   int row_offset, void* u, void* v, float* inv_norm_img, float* inv_norm_txt,
                   float* logit_scale_out, float* S, float* row_lse, float* col_max, float* col_sum, float* diag);
int forde_clip_col_rescale(void* stream, const float* col_max_local, const float* col_max_global, float* col_sum,
                           int Bg);
int forde_clip_loss(void* stream, const float* row_lse, const float* diag, const float* col_max,
                    const float* col_sum, int Bl, int Bg, int row_offset, float* col_lse, float* loss);
/* bwd:  dS = (softmax_row + softmax_col)/(2 Bg) - onehot/Bg  (bf16 scratch [Bl][round_up(Bg, 8)]); col_lse is the GLOBAL
 *       column log-sum-exp; logit_scale is the device scalar written by forde_clip_fwd;
 *       dU [Bl][E] = exp(tau) * dS . v (fp32, complete), dV [Bg][E] = exp(tau) * dS^T . u (fp32, partial over
 *       ranks: reduce-scatter it), d_log_temp[1] += sum dS * S (may be NULL; all-reduce(SUM) across ranks).
 *       Follow with forde_l2norm_bwd to go from dU / dV to the gradients of the un-normalised embeddings. */
int forde_clip_bwd(void* stream, const void* u, const void* v, const float* S, const float* row_lse,
                   const float* col_lse, const float* logit_scale, int Bl, int Bg, int E, int row_offset, void* dS,
                   float* dU, float* dV, float* d_log_temp);
/* normalisation backward for rows of an embedding: d_x = inv_norm * (dU - u (u . dU)) */
int forde_l2norm_bwd(void* stream, const void* u, const float* inv_norm, const float* dU, float* d_x, int B,
                     int E);

/* ---- MoE variant of the FFN slot (BASELINE config 4) --------------------------------------------------------
 * Reference: src/forde/moe.py:57-281 (MoELayer / MoEStatefulLayer).  Same result as the reference's "every expert on
 * every token, then gather the top-k" (moe.py:163,169-185), computed on the routed tokens only.  All routing state
 * stays on the device (no host synchronisation): the caller allocates
 *   R = forde_moe_capacity_rows(M, k, E) rows for every "gathered" buffer (xg, zg, hg, yg, dyg, dzg, dxg);
 *   logits, probs [M][E] f32 ; topk_idx [M][k] i32 ; topk_w [M][k] f32 ; counts [E] i32 ;
 *   prob_partials [forde_moe_num_prob_partials(M)][E] f32 ; seg_offsets [E+1] i32 ; n_blocks_dev [1] i32 ;
 *   block_expert [R/128] i32 ; slot_of [M*k] i32 ; src_of_slot [R] i32.
 * route:   logits = x . w_r + b_r (w_r [d][E] bf16; moe.py:88), probs = softmax (moe.py:90), top-k with
 *          jax.lax.top_k order (descending, equal values -> lower index; moe.py:130), topk_w = softmax of the k
 *          logits (moe.py:133), counts = bincount(topk_idx) (moe.py:216, exact), aux_loss[1] = aux_weight * E *
 *          sum_e counts_e/(M k) * mean_m probs_e (moe.py:212-222), expert_usage[E] += mean_m probs (moe.py:277-278),
 *          step_count += 1 (moe.py:279); then the routing plan: expert e owns rows seg_offsets[e]..seg_offsets[e+1]
 *          (128-aligned) of the gathered buffers, entry (token, j) sits in row slot_of[token*k + j], rows keep token
 *          order inside an expert, src_of_slot is the inverse map (-1 for padding rows).
 * gather / experts_fwd / combine: xg[slot] = x[token]; zg = xg . w_up[e] + b_up[e]; hg = gelu_tanh(zg)
 *          (flax nn.gelu default, moe.py:27); yg = hg . w_down[e] + b_down[e]; out[token] = sum_j topk_w * yg[slot].
 *          Expert weights are stacked: w_up_all [E][d][H], w_down_all [E][H][d] bf16, biases [E][H], [E][d] f32.
 * backward: scatter (dyg[slot] = topk_w * dout[token]), gate_grad (d_topk_w = <dout[token], yg[slot]>),
 *          experts_bwd (dzg, dxg, and per expert dw_up_all [E][d][H] (written), dw_down_all [E][H][d] (written),
 *          db_up_all [E][H] +=, db_down_all [E][d] +=), then combine(dxg, slot_of, NULL) sums the k copies into dx. */
int forde_moe_capacity_rows(int M, int k, int E);
int forde_moe_num_prob_partials(int M);
int forde_moe_plan_scratch_ints(int M, int k, int E); /* ints of plan_scratch: per-chunk, per-expert entry counts */
int forde_moe_route(void* stream, const void* x, const void* w_r, const float* b_r, int M, int d, int E, int k,
                    float aux_weight, float* logits, float* probs, int* topk_idx, float* topk_w, int* counts,
                    float* prob_partials, int* plan_scratch, int* seg_offsets, int* n_blocks_dev, int* block_expert,
                    int* slot_of, int* src_of_slot, float* aux_loss, float* expert_usage, int* step_count);
int forde_moe_gather(void* stream, const void* x, const int* src_of_slot, int k, void* xg, int rows, int d);
int forde_moe_scatter(void* stream, const void* dout, const int* src_of_slot, const float* topk_w, int k, void* dyg,
                      int rows, int d);
int forde_moe_combine(void* stream, const void* yg, const int* slot_of, const float* topk_w, void* out, int M, int k,
                      int d);
int forde_moe_gate_grad(void* stream, const void* dout, const void* yg, const int* slot_of, float* d_topk_w, int M,
                        int k, int d);
/* router backward: dlogits from d_topk_w (through the top-k softmax) and from d(probs) = d_aux[0] * d(aux)/d(probs)
 * (+ optional d_probs [M][E]); then dx[M][d] (bf16) += dlogits . w_r^T IN PLACE, dw_r [d][E] += x^T . dlogits,
 * db_r [E] += sum_m dlogits (fp32 atomics).  d_aux is a device scalar (may be NULL = 0). */
int forde_moe_router_bwd(void* stream, const void* x, const void* w_r, const float* probs, const int* topk_idx,
                         const float* topk_w, const float* d_topk_w, const int* counts, const float* d_aux,
                         const float* d_probs, float aux_weight, int M, int d, int E, int k, void* dx, float* dw_r,
                         float* db_r);
int forde_moe_experts_fwd(void* stream, const void* xg, const void* w_up_all, const float* b_up_all,
                          const void* w_down_all, const float* b_down_all, const int* n_blocks_dev,
                          const int* block_expert, void* zg, void* hg, void* yg, int r_cap, int d, int H, int E);
int forde_moe_experts_bwd(void* stream, const void* xg, const void* dyg, const void* w_up_all,
                          const void* w_down_all, const void* zg, const void* hg, const int* n_blocks_dev,
                          const int* block_expert, const int* seg_offsets, void* dzg, void* dxg, float* dw_up_all,
                          float* db_up_all, float* dw_down_all, float* db_down_all, int r_cap, int d, int H, int E);

/* ---- fused residual-add + LayerNorm: the `add/norm` either side of the StatefulLayer in a FORDE block
 *      (README.md:42-43; flax nn.LayerNorm, eps 1e-6).  Caller side of the path. -------
 * fwd: out = LayerNorm(x + r) * gamma + beta  (x, r, out bf16 [rows][cols]; r may be NULL; gamma, beta fp32);
 *      mean[rows], rstd[rows] fp32 are saved for the backward.  cols % 8 == 0, cols <= 1024.
 * bwd: ds = d(x + r) (the gradient of BOTH x and r), d_gamma[cols] += , d_beta[cols] += ;
 *      partials is scratch of forde_ln_num_partials(rows) * 2 * cols floats. */
int forde_ln_num_partials(int rows);
int forde_add_layernorm_fwd(void* stream, const void* x, const void* r, const float* gamma, const float* beta,
                            void* out, float* mean, float* rstd, int rows, int cols, float eps);
int forde_add_layernorm_bwd(void* stream, const void* dy, const void* x, const void* r, const float* gamma,
                            const float* mean, const float* rstd, void* ds, float* partials, float* d_gamma,
                            float* d_beta, int rows, int cols);

/* ---- optimizer step of the fast loop ---------------------------------------------------------
 * Reference: optax.chain(optax.clip_by_global_norm(1.0), optax.adamw(lr, weight_decay)) built at
 * src/training/train.py:74-77 and applied by state.apply_gradients at :162; the monitored grad_norm of :168.
 * All parameters / gradients / moments live in flat, 64-element-padded fp32 buffers of n elements (n % 4 == 0,
 * 16-byte aligned); pad elements are zero and stay zero.  Exactly one of g32 / g16 is non-NULL: fp32 gradients, or
 * the bf16 gradients that come back from a bf16 all-reduce.
 *   forde_grad_sqnorm     : partials[forde_optim_num_partials(n)] = per-block sums of g^2 (fixed order, so the
 *                           norm is bit-reproducible); *step_count += 1 (the optax count; may be NULL).
 *   forde_adamw_clip_step : norm = sqrt(sum partials); g *= (norm < max_norm ? 1 : max_norm / norm)   [no epsilon,
 *                           optax.clip_by_global_norm]; mu = b1 mu + (1-b1) g; nu = b2 nu + (1-b2) g^2;
 *                           p -= lr * (mu/(1-b1^t) / (sqrt(nu/(1-b2^t)) + eps) + wd * p), t = *step_count (already
 *                           incremented by forde_grad_sqnorm); p16 (may be NULL) = bf16(p), the compute mirror;
 *                           grad_norm_out[1] (may be NULL) = the pre-clip norm.  max_norm <= 0 disables clipping.
 *                           Hyper-parameters are doubles (Python floats in optax): 1 - b is formed in double before
 *                           it is rounded to fp32, exactly as optax's weakly typed scalars are.
 *   forde_cast_f32_bf16   : dst = bf16(src) (gradient buckets before a bf16 all-reduce). */
int forde_optim_num_partials(long long n);
int forde_grad_sqnorm(void* stream, const float* g32, const void* g16, long long n, float* partials, int* step_count);
int forde_adamw_clip_step(void* stream, float* p32, const float* g32, const void* g16, float* mu, float* nu, void* p16,
                          long long n, const float* partials, int num_partials, double max_norm, double lr, double b1,
                          double b2, double eps, double wd, const int* step_count, float* grad_norm_out);
int forde_cast_f32_bf16(void* stream, const float* src, void* dst, long long n);

/* ---- front ends of the two towers (README.md:30-34; layouts of src/data/download_to_drive.py:106-108,
 *      171-178: image (S,S,3) NHWC, input_ids int32).  HBM-bound gathers / scatters; the patch-embedding and projection
 *      head GEMMs run on forde_gemm_bf16 / forde_gemm_f32.  Gradients are accumulated (fp32 red.add) into the caller's
 *      buffers like every other weight gradient.
 *   patchify:            patches[b*g*g + gy*g + gx][(py*p + px)*3 + ch] = images[b][gy*p+py][gx*p+px][ch], g = S/p;
 *                        ld = output row pitch in elements (>= p*p*3; round it up to 8 for the GEMM)
 *   vision_assemble_fwd: x[b][0] = cls + pos[0]; x[b][1+i] = patch_tok[b*G + i] + pos[1+i]       (bf16, w % 8 == 0)
 *   vision_assemble_bwd: d_patch_tok[b*G+i] = dx[b][1+i]; d_pos[t] += sum_b dx[b][t]; d_cls += sum_b dx[b][0]
 *                        (d_patch_tok, d_cls, d_pos may each be NULL)
 *   embed_tokens_fwd:    x[b][t] = table[ids[b][t]] + pos[t]   (ids clamped into [0, V))
 *   embed_tokens_bwd:    d_table[ids[b][t]] += dx[b][t]; d_pos[t] += sum_b dx[b][t]   (d_pos may be NULL) */
int forde_patchify(void* stream, const void* images, void* patches, long long ld, int B, int S, int p);
int forde_vision_assemble_fwd(void* stream, const void* patch_tok, const void* cls, const void* pos, void* x, int B,
                              int G, int w);
int forde_vision_assemble_bwd(void* stream, const void* dx, void* d_patch_tok, float* d_cls, float* d_pos, int B, int G,
                              int w);
int forde_embed_tokens_fwd(void* stream, const int* ids, const void* table, const void* pos, void* x, int B, int T,
                           int w, int V);
int forde_embed_tokens_bwd(void* stream, const int* ids, const void* dx, float* d_table, float* d_pos, int B, int T,
                           int w, int V);

/* ---- self-attention core for short sequences (README.md:42-43 "self-attention -> add/norm") -------
 * softmax(q k^T / sqrt(64)) v per (batch, head), head_dim = 64, T <= forde_attn_max_tokens() (= 128: the text tower's
 * 77 tokens): the whole sequence of a head is ONE tcgen05 tile (scores in TMEM, probabilities staged once in shared
 * memory and consumed K-major and MN-major).  Longer sequences: forde_attn2_* below.
 * q, k, v, o, d_o, dq, dk, dv: bf16 [B*T][ld], head h in columns [64 h, 64 h + 64); 16-byte aligned, ld % 8 == 0.
 * lse [B][heads][T] fp32 = ln sum_j exp(q.k_j / 8), written by fwd, read by bwd. */
int forde_attn_max_tokens(void);
int forde_attn_fwd(void* stream, const void* q, const void* k, const void* v, void* o, float* lse, int B, int T,
                   int heads, long long ld);
int forde_attn_bwd(void* stream, const void* q, const void* k, const void* v, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int T, int heads, long long ld);

/* ---- the same for medium sequences, 1 <= T <= forde_attn2_max_tokens() (= 256: the ViT-B/16 vision tower's 197
 * tokens).  One persistent CTA per SM walks the (batch, head) items: all operands of an item in shared
 * memory (next item's loads in flight), scores of both 128-row blocks in TMEM.  Forward: one warp group per row
 * block, a thread owns a score row, row block 1 issued half an item behind row block 0, o written by TMA (so `o` must
 * be a [B][T][ld] tensor the 3-D tensor map can describe: 16-byte aligned, ld % 8 == 0).  Backward: 8 softmax warps
 * (TMEM lane quarter x column half) + 4 read-out warps; it additionally reads the forward output `o`
 * (D = rowsum(dO * o)). */
int forde_attn2_max_tokens(void);
int forde_attn2_fwd(void* stream, const void* q, const void* k, const void* v, void* o, float* lse, int B, int T,
                    int heads, long long ld);
int forde_attn2_bwd(void* stream, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                    const float* lse, void* dq, void* dk, void* dv, int B, int T, int heads, long long ld);

#ifdef __cplusplus
}
#endif
#endif /* FORDE_B200_H_ */
