// Front ends of the two towers (README.md:30-34: "patchify -> +CLS, +pos -> blocks" and
// "embed -> +pos -> blocks"; tensor layouts from src/data/download_to_drive.py:106-108,171-178: image (S,S,3) NHWC,
// input_ids int32).  All HBM-bound gather / scatter kernels; the patch-embedding and projection-head GEMMs themselves
// run on the tcgen05 engine (forde_gemm_bf16 / forde_gemm_f32), so the front ends are:
//   forde_patchify            images [B][S][S][3] -> patch rows [B*g*g][p*p*3]   (pure layout change, 16-byte moves)
//   forde_vision_assemble_fwd x[b][0] = cls + pos[0] ; x[b][1+i] = patch_tok[b*G+i] + pos[1+i]
//   forde_vision_assemble_bwd d_patch_tok[b*G+i] = dx[b][1+i] ; d_pos[t] += sum_b dx[b][t] ; d_cls += sum_b dx[b][0]
//   forde_embed_tokens_fwd    x[b][t] = table[ids[b][t]] + pos[t]
//   forde_embed_tokens_bwd    d_table[ids[b][t]] += dx[b][t] ; d_pos[t] += sum_b dx[b][t]      (fp32 red.add)
// Gradients are ACCUMULATED into fp32 buffers (the train state's flat gradient views), like every other weight
// gradient of the path.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/forde_b200.h"
#include "forde_host.h"

namespace forde {

__device__ __forceinline__ void unpack8f(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// One thread moves one 16-byte piece (8 bf16).  A patch row is p rows of the image, each p*3 contiguous values.
__global__ void patchify_kernel(const uint4* __restrict__ img, uint4* __restrict__ out, int B, int S, int p) {
  const int g = S / p;
  const int vec_per_prow = p * 3 / 8;                   // 16-byte pieces per patch pixel-row
  const long long total = static_cast<long long>(B) * g * g * p * vec_per_prow;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_prow);
    long long r = i / vec_per_prow;
    const int py = static_cast<int>(r % p); r /= p;
    const int gx = static_cast<int>(r % g); r /= g;
    const int gy = static_cast<int>(r % g);
    const long long b = r / g;
    // source: image b, pixel row gy*p+py, pixel column gx*p, 3 channels
    const long long src = ((b * S + gy * p + py) * S + gx * p) * 3 / 8 + v;
    out[i] = __ldg(img + src);
  }
}

// element-wise variant for patch sizes whose pixel rows are not whole 16-byte pieces (p = 14: 42 values); `ld` is the
// output row pitch (round_up(p*p*3, 8) so the rows stay TMA-addressable; the pad columns are never read: the GEMM's
// tensor map ends at p*p*3 and TMA zero-fills beyond it)
__global__ void patchify_scalar_kernel(const uint16_t* __restrict__ img, uint16_t* __restrict__ out, int B, int S,
                                       int p, int ld) {
  const int g = S / p, row_len = p * p * 3, prow = p * 3;
  const long long total = static_cast<long long>(B) * g * g * row_len;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % row_len);
    long long r = i / row_len;                       // patch row index b*g*g + gy*g + gx
    const int gx = static_cast<int>(r % g);
    const int gy = static_cast<int>((r / g) % g);
    const long long b = r / (static_cast<long long>(g) * g);
    const int py = e / prow, rest = e - py * prow;
    out[r * ld + e] = img[((b * S + gy * p + py) * S + gx * p) * 3 + rest];
  }
}

// grid.x = token t (0..G), grid.y = batch chunk; block = w/8 threads (one 16-byte column group each)
__global__ void vision_assemble_fwd_kernel(const uint4* __restrict__ tok, const uint4* __restrict__ cls,
                                           const uint4* __restrict__ pos, uint4* __restrict__ x, int B, int G, int wv,
                                           int b_per_block) {
  const int t = blockIdx.x, c = threadIdx.x;
  if (c >= wv) return;
  float pv[8];
  unpack8f(__ldg(pos + static_cast<size_t>(t) * wv + c), pv);
  if (t == 0) {
    float cv[8];
    unpack8f(__ldg(cls + c), cv);
#pragma unroll
    for (int k = 0; k < 8; ++k) pv[k] += cv[k];
  }
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    float o[8];
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = pv[k];
    } else {
      unpack8f(__ldg(tok + (static_cast<size_t>(b) * G + (t - 1)) * wv + c), o);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += pv[k];
    }
    x[(static_cast<size_t>(b) * (G + 1) + t) * wv + c] = pack8f(o);
  }
}

__global__ void vision_assemble_bwd_kernel(const uint4* __restrict__ dx, uint4* __restrict__ d_tok,
                                           float* __restrict__ d_cls, float* __restrict__ d_pos, int B, int G, int wv,
                                           int b_per_block) {
  const int t = blockIdx.x, c = threadIdx.x;
  if (c >= wv) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    const uint4 q = __ldg(dx + (static_cast<size_t>(b) * (G + 1) + t) * wv + c);
    if (t > 0 && d_tok != nullptr) d_tok[(static_cast<size_t>(b) * G + (t - 1)) * wv + c] = q;
    float f[8];
    unpack8f(q, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  float* dp = d_pos + (static_cast<size_t>(t) * wv + c) * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (d_pos != nullptr) atomicAdd(dp + k, acc[k]);
    if (t == 0 && d_cls != nullptr) atomicAdd(d_cls + c * 8 + k, acc[k]);
  }
}

// grid.x = position t, grid.y = batch chunk; block = w/8 threads
__global__ void embed_tokens_fwd_kernel(const int* __restrict__ ids, const uint4* __restrict__ table,
                                        const uint4* __restrict__ pos, uint4* __restrict__ x, int B, int T, int wv, int V,
                                        int b_per_block) {
  const int t = blockIdx.x, c = threadIdx.x;
  if (c >= wv) return;
  float pv[8];
  unpack8f(__ldg(pos + static_cast<size_t>(t) * wv + c), pv);
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    int id = ids[static_cast<size_t>(b) * T + t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);           // clamp like jnp's default gather mode for out-of-range ids
    float o[8];
    unpack8f(__ldg(table + static_cast<size_t>(id) * wv + c), o);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] += pv[k];
    x[(static_cast<size_t>(b) * T + t) * wv + c] = pack8f(o);
  }
}

__global__ void embed_tokens_bwd_kernel(const int* __restrict__ ids, const uint4* __restrict__ dx,
                                        float* __restrict__ d_table, float* __restrict__ d_pos, int B, int T, int wv,
                                        int V, int b_per_block) {
  const int t = blockIdx.x, c = threadIdx.x;
  if (c >= wv) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    int id = ids[static_cast<size_t>(b) * T + t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float f[8];
    unpack8f(__ldg(dx + (static_cast<size_t>(b) * T + t) * wv + c), f);
    float* dt = d_table + (static_cast<size_t>(id) * wv + c) * 8;
    // two 16-byte vector reductions per piece (sm_90+: red.global.add.v4.f32)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3])
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]),
                 "f"(f[7])
                 : "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  if (d_pos != nullptr) {
    float* dp = d_pos + (static_cast<size_t>(t) * wv + c) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(dp + k, acc[k]);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// batch rows per block so that tokens x chunks is a few waves of the machine
static int batch_chunk(int B, int tokens) {
  const int sms = sm_count();
  if (sms <= 0) return 0;
  int chunks = (8 * sms + tokens - 1) / tokens;
  if (chunks < 1) chunks = 1;
  if (chunks > B) chunks = B;
  return (B + chunks - 1) / chunks;
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_patchify(void* stream, const void* images, void* patches, long long ld, int B, int S, int p) {
  if (images == nullptr || patches == nullptr || B <= 0 || S <= 0 || p <= 0 || S % p != 0 || ld < p * p * 3)
    return FORDE_EINVAL;
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  if ((p * 3) % 8 == 0 && ld == p * p * 3 && aligned16(images) && aligned16(patches)) {
    patchify_kernel<<<8 * sms, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(images), static_cast<uint4*>(patches), B, S, p);
  } else {
    patchify_scalar_kernel<<<8 * sms, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(images), static_cast<uint16_t*>(patches), B, S, p, static_cast<int>(ld));
  }
  return check_launch("patchify_kernel");
}

int forde_vision_assemble_fwd(void* stream, const void* patch_tok, const void* cls, const void* pos, void* x, int B,
                              int G, int w) {
  if (patch_tok == nullptr || cls == nullptr || pos == nullptr || x == nullptr || B <= 0 || G <= 0 || w <= 0)
    return FORDE_EINVAL;
  if ((w & 7) != 0 || w > 8192 || !aligned16(patch_tok) || !aligned16(cls) || !aligned16(pos) || !aligned16(x))
    return FORDE_EALIGN;
  const int bpb = batch_chunk(B, G + 1);
  if (bpb <= 0) return FORDE_ENODEV;
  vision_assemble_fwd_kernel<<<dim3(G + 1, (B + bpb - 1) / bpb), (w / 8 + 31) / 32 * 32, 0,
                               static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(patch_tok), static_cast<const uint4*>(cls), static_cast<const uint4*>(pos),
      static_cast<uint4*>(x), B, G, w / 8, bpb);
  return check_launch("vision_assemble_fwd_kernel");
}

int forde_vision_assemble_bwd(void* stream, const void* dx, void* d_patch_tok, float* d_cls, float* d_pos, int B, int G,
                              int w) {
  if (dx == nullptr || B <= 0 || G <= 0 || w <= 0) return FORDE_EINVAL;
  if ((w & 7) != 0 || w > 8192 || !aligned16(dx) || !aligned16(d_patch_tok)) return FORDE_EALIGN;
  const int bpb = batch_chunk(B, G + 1);
  if (bpb <= 0) return FORDE_ENODEV;
  vision_assemble_bwd_kernel<<<dim3(G + 1, (B + bpb - 1) / bpb), (w / 8 + 31) / 32 * 32, 0,
                               static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dx), static_cast<uint4*>(d_patch_tok), d_cls, d_pos, B, G, w / 8, bpb);
  return check_launch("vision_assemble_bwd_kernel");
}

int forde_embed_tokens_fwd(void* stream, const int* ids, const void* table, const void* pos, void* x, int B, int T,
                           int w, int V) {
  if (ids == nullptr || table == nullptr || pos == nullptr || x == nullptr || B <= 0 || T <= 0 || w <= 0 || V <= 0)
    return FORDE_EINVAL;
  if ((w & 7) != 0 || w > 8192 || !aligned16(table) || !aligned16(pos) || !aligned16(x)) return FORDE_EALIGN;
  const int bpb = batch_chunk(B, T);
  if (bpb <= 0) return FORDE_ENODEV;
  embed_tokens_fwd_kernel<<<dim3(T, (B + bpb - 1) / bpb), (w / 8 + 31) / 32 * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<const uint4*>(table), static_cast<const uint4*>(pos), static_cast<uint4*>(x), B, T, w / 8, V, bpb);
  return check_launch("embed_tokens_fwd_kernel");
}

int forde_embed_tokens_bwd(void* stream, const int* ids, const void* dx, float* d_table, float* d_pos, int B, int T,
                           int w, int V) {
  if (ids == nullptr || dx == nullptr || d_table == nullptr || B <= 0 || T <= 0 || w <= 0 || V <= 0)
    return FORDE_EINVAL;
  if ((w & 7) != 0 || w > 8192 || !aligned16(dx) || !aligned16(d_table)) return FORDE_EALIGN;
  const int bpb = batch_chunk(B, T);
  if (bpb <= 0) return FORDE_ENODEV;
  embed_tokens_bwd_kernel<<<dim3(T, (B + bpb - 1) / bpb), (w / 8 + 31) / 32 * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<const uint4*>(dx), d_table, d_pos, B, T, w / 8, V, bpb);
  return check_launch("embed_tokens_bwd_kernel");
}

}  // extern "C"
