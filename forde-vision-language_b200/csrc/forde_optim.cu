// Optimizer step of the fast loop as two HBM-streaming passes over the flat parameter buffers
// (reference: optax.chain(clip_by_global_norm(1.0), adamw(lr, weight_decay)) in
// /root/reference/src/training/train.py:74-77, applied by state.apply_gradients at :162, and the monitoring
// grad_norm at :168):
//
//   pass 1  forde_grad_sqnorm      : per-block sums of g^2 (fixed-order tree per block -> bit-reproducible), and the
//                                    device step counter += 1 (stream-ordered before pass 2 reads it);
//   pass 2  forde_adamw_clip_step  : every block re-reduces the (<= 2368) block sums in the same fixed order in
//                                    double -> the identical global norm; clip scale = max_norm / norm if
//                                    norm >= max_norm else 1 (optax.clip_by_global_norm: no epsilon);
//                                    mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ; bias-corrected
//                                    u = mu_hat / (sqrt(nu_hat) + eps) + wd * p ; p -= lr * u ; bf16 mirror = bf16(p).
//
// Bytes per parameter: pass 1 reads 4 (fp32 grads) or 2 (bf16 grads after a bf16 all-reduce); pass 2 reads
// g + p + mu + nu (16 or 14) and writes p + mu + nu + bf16 p (14): 34 B/param against 50 B/param for the
// norm / scale / fused-AdamW / cast sequence of library calls it replaces.  No host synchronisation anywhere.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/forde_b200.h"
#include "forde_host.h"

namespace forde {

constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCKS_PER_SM = 8;
constexpr int OPT_MAX_PARTIALS = 4096;

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (wid == 0) {
    t = lane < (OPT_THREADS >> 5) ? sm[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}

__device__ __forceinline__ float4 load_grad4(const float* g32, const __nv_bfloat16* g16, long long i) {
  if (g32 != nullptr) return *reinterpret_cast<const float4*>(g32 + i);
  const uint2 q = *reinterpret_cast<const uint2*>(g16 + i);
  return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u), __uint_as_float(q.y << 16),
                     __uint_as_float(q.y & 0xFFFF0000u));
}

// n is a multiple of 4 (the flat buffers are padded to 64 elements per leaf)
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const float* __restrict__ g32, const __nv_bfloat16* __restrict__ g16, long long n,
                   float* __restrict__ partials, int* step_count) {
  __shared__ float sm[32];
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * OPT_THREADS * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * OPT_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const float4 g = load_grad4(g32, g16, i);
    acc = fmaf(g.x, g.x, acc);
    acc = fmaf(g.y, g.y, acc);
    acc = fmaf(g.z, g.z, acc);
    acc = fmaf(g.w, g.w, acc);
  }
  const float t = block_sum_256(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = t;
    if (blockIdx.x == 0 && step_count != nullptr) *step_count += 1;
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_clip_kernel(float* __restrict__ p32, const float* __restrict__ g32, const __nv_bfloat16* __restrict__ g16,
                  float* __restrict__ mu, float* __restrict__ nu, __nv_bfloat16* __restrict__ p16, long long n,
                  const float* __restrict__ partials, int num_partials, float max_norm, float lr, float b1, float b2,
                  float omb1, float omb2, float eps, float wd, const int* __restrict__ step_count,
                  float* __restrict__ grad_norm_out) {
  __shared__ double red[OPT_THREADS];
  __shared__ float s_scale, s_c1, s_c2;
  // the same fixed-order reduction in every block -> every block sees the identical norm
  double a = 0.0;
  for (int i = threadIdx.x; i < num_partials; i += OPT_THREADS) a += static_cast<double>(partials[i]);
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = OPT_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = static_cast<float>(sqrt(red[0]));
    s_scale = (norm < max_norm || max_norm <= 0.f) ? 1.f : max_norm / norm;  // optax.clip_by_global_norm
    const float t = static_cast<float>(*step_count);                          // already incremented by pass 1
    s_c1 = 1.f / (1.f - powf(b1, t));
    s_c2 = 1.f / (1.f - powf(b2, t));
    if (blockIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = norm;   // pre-clip norm (train.py:168)
  }
  __syncthreads();
  const float scale = s_scale, c1 = s_c1, c2 = s_c2;
  const long long stride = static_cast<long long>(gridDim.x) * OPT_THREADS * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * OPT_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const float4 g4 = load_grad4(g32, g16, i);
    float4 p = *reinterpret_cast<const float4*>(p32 + i);
    float4 m = *reinterpret_cast<const float4*>(mu + i);
    float4 v = *reinterpret_cast<const float4*>(nu + i);
    const float g[4] = {g4.x * scale, g4.y * scale, g4.z * scale, g4.w * scale};
    float* pp = reinterpret_cast<float*>(&p);
    float* mm = reinterpret_cast<float*>(&m);
    float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] = b1 * mm[k] + omb1 * g[k];            // omb = 1 - b evaluated in double on the host, as optax's
      vv[k] = b2 * vv[k] + omb2 * g[k] * g[k];     // Python-float (1 - decay) is before it meets the fp32 arrays
      const float u = (mm[k] * c1) / (sqrtf(vv[k] * c2) + eps) + wd * pp[k];
      pp[k] = pp[k] - lr * u;
    }
    *reinterpret_cast<float4*>(p32 + i) = p;
    *reinterpret_cast<float4*>(mu + i) = m;
    *reinterpret_cast<float4*>(nu + i) = v;
    if (p16 != nullptr) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p16 + i) = o;
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * OPT_THREADS * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * OPT_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const float4 p = *reinterpret_cast<const float4*>(src + i);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo);
    o.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
}

static int opt_blocks(long long n) {
  const int sms = sm_count();
  if (sms <= 0) return 0;
  long long want = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  long long cap = static_cast<long long>(sms) * OPT_BLOCKS_PER_SM;
  if (cap > OPT_MAX_PARTIALS) cap = OPT_MAX_PARTIALS;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_optim_num_partials(long long n) {
  if (n <= 0) return 0;
  const int b = opt_blocks(n);
  return b > 0 ? b : OPT_MAX_PARTIALS;  // without a device: an upper bound, so buffers can be sized on the host
}

int forde_grad_sqnorm(void* stream, const float* g32, const void* g16, long long n, float* partials, int* step_count) {
  if ((g32 == nullptr) == (g16 == nullptr) || partials == nullptr || n <= 0) return FORDE_EINVAL;
  if ((n & 3) != 0 || (reinterpret_cast<uintptr_t>(g32) & 15) != 0 || (reinterpret_cast<uintptr_t>(g16) & 7) != 0)
    return FORDE_EALIGN;
  const int blocks = opt_blocks(n);
  if (blocks <= 0) return FORDE_ENODEV;
  grad_sqnorm_kernel<<<blocks, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      g32, static_cast<const __nv_bfloat16*>(g16), n, partials, step_count);
  return check_launch("grad_sqnorm_kernel");
}

int forde_adamw_clip_step(void* stream, float* p32, const float* g32, const void* g16, float* mu, float* nu, void* p16,
                          long long n, const float* partials, int num_partials, double max_norm, double lr, double b1,
                          double b2, double eps, double wd, const int* step_count, float* grad_norm_out) {
  if (p32 == nullptr || (g32 == nullptr) == (g16 == nullptr) || mu == nullptr || nu == nullptr ||
      partials == nullptr || step_count == nullptr || n <= 0 || num_partials <= 0)
    return FORDE_EINVAL;
  if ((n & 3) != 0 || (reinterpret_cast<uintptr_t>(p32) & 15) != 0 || (reinterpret_cast<uintptr_t>(g32) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(g16) & 7) != 0 || (reinterpret_cast<uintptr_t>(mu) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(nu) & 15) != 0 || (reinterpret_cast<uintptr_t>(p16) & 7) != 0)
    return FORDE_EALIGN;
  const int blocks = opt_blocks(n);
  if (blocks <= 0) return FORDE_ENODEV;
  adamw_clip_kernel<<<blocks, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      p32, g32, static_cast<const __nv_bfloat16*>(g16), mu, nu, static_cast<__nv_bfloat16*>(p16), n, partials,
      num_partials, static_cast<float>(max_norm), static_cast<float>(lr), static_cast<float>(b1),
      static_cast<float>(b2), static_cast<float>(1.0 - b1), static_cast<float>(1.0 - b2), static_cast<float>(eps),
      static_cast<float>(wd), step_count, grad_norm_out);
  return check_launch("adamw_clip_kernel");
}

int forde_cast_f32_bf16(void* stream, const float* src, void* dst, long long n) {
  if (src == nullptr || dst == nullptr || n <= 0) return FORDE_EINVAL;
  if ((n & 3) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 7) != 0)
    return FORDE_EALIGN;
  const int blocks = opt_blocks(n);
  if (blocks <= 0) return FORDE_ENODEV;
  cast_f32_bf16_kernel<<<blocks, OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_f32_bf16_kernel");
}

}  // extern "C"
