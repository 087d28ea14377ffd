// C-ABI entry points for the tcgen05 GEMM family (see include/forde_b200.h).
// Host side only: argument validation, TMA tensor-map construction (cached), launch configuration.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/forde_b200.h"
#include "forde_gemm.cuh"
#include "forde_host.h"

namespace forde {

thread_local char g_last_err[256] = {0};

void record_cuda_error(const char* what, const char* msg) {
  std::snprintf(g_last_err, sizeof(g_last_err), "%s: %s", what, msg);
}

static std::atomic<unsigned long long> g_launches{0};

int check_launch(const char* what, int kernels) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    record_cuda_error(what, cudaGetErrorString(e));
    return FORDE_ECUDA;
  }
  g_launches.fetch_add(static_cast<unsigned long long>(kernels), std::memory_order_relaxed);
  return FORDE_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  int dev;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && dev == o.dev;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer); mix(static_cast<uint64_t>(k.dev));
    return h;
  }
};

// 2-D row-major bf16 tensor [outer][inner] with row pitch `ld` elements, 128B swizzle, zero OOB fill.
// The encoded map depends only on the key, so it is cached (one-time descriptor cache, mutex guarded).
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                   uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0) return FORDE_EALIGN;
  int dev = 0;
  cudaGetDevice(&dev);
  TmapKey key{ptr, inner, outer, ld, box_inner, box_outer, dev};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      std::memcpy(out, &it->second, sizeof(CUtensorMap));
      return FORDE_OK;
    }
  }
  // cuTensorMapEncodeTiled is a driver call and needs a current context; a thread that has only ever had
  // cudaSetDevice() called on it (torch's autograd worker threads) may not have the primary context bound yet.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    record_cuda_error("cuTensorMapEncodeTiled", "driver entry point not available");
    return FORDE_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "CUresult %d", static_cast<int>(r));
    record_cuda_error("cuTensorMapEncodeTiled", buf);
    return FORDE_ECUDA;
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  std::memcpy(out, &m, sizeof(CUtensorMap));
  return FORDE_OK;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous; pitches ld1, ld2 in elements), box {b0, b1, 1}, 128B swizzle, zero OOB
// fill: the per-(batch, head) token tiles of the attention kernels (rows beyond the sequence read as zeros).
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                      uint64_t ld2, uint32_t b0, uint32_t b1) {
  static std::mutex mu;
  struct Key3 {
    const void* ptr;
    uint64_t d0, d1, d2, ld1, ld2;
    uint32_t b0, b1;
    int dev;
    bool operator==(const Key3& o) const {
      return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld1 == o.ld1 && ld2 == o.ld2 && b0 == o.b0 &&
             b1 == o.b1 && dev == o.dev;
    }
  };
  struct Key3Hash {
    size_t operator()(const Key3& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr);
      auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
      mix(k.d0); mix(k.d1); mix(k.d2); mix(k.ld1); mix(k.ld2); mix(k.b0); mix(k.b1); mix(static_cast<uint64_t>(k.dev));
      return h;
    }
  };
  static std::unordered_map<Key3, CUtensorMap, Key3Hash> cache;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld1 % 8) != 0 || (ld2 % 8) != 0) return FORDE_EALIGN;
  int dev = 0;
  cudaGetDevice(&dev);
  Key3 key{ptr, d0, d1, d2, ld1, ld2, b0, b1, dev};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      std::memcpy(out, &it->second, sizeof(CUtensorMap));
      return FORDE_OK;
    }
  }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    record_cuda_error("cuTensorMapEncodeTiled", "driver entry point not available");
    return FORDE_ECUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "CUresult %d (3-D map)", static_cast<int>(r));
    record_cuda_error("cuTensorMapEncodeTiled", buf);
    return FORDE_ECUDA;
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  std::memcpy(out, &m, sizeof(CUtensorMap));
  return FORDE_OK;
}

// operand map: `rows` is the M (or N) extent, `k` the reduction extent.
//   K-major : memory [rows][k]  -> dims {k, rows},  box {64, tile_rows}
//   MN-major: memory [k][rows]  -> dims {rows, k},  box {64, 64}
static int make_operand_map(CUtensorMap* out, const void* ptr, bool mn_major, long long ld, int rows, int k,
                            int tile_rows) {
  if (mn_major) return make_tmap_bf16(out, ptr, rows, k, ld, 64, BK);
  return make_tmap_bf16(out, ptr, k, rows, ld, BK, tile_rows);
}

template <bool A_MN, bool B_MN, int EPI, bool PAIR>
static int launch_inst(const GemmParams& p, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  constexpr int smem = gemm_smem_bytes(EPI, PAIR);
  // the attribute is per device function (per context); set once per process/device is enough for one GPU
  // per process, which is the deployment model (one process per GPU).
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(forde_gemm_kernel<A_MN, B_MN, EPI, PAIR>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) {
    record_cuda_error("cudaFuncSetAttribute", cudaGetErrorString(attr_err));
    return FORDE_ECUDA;
  }
  const int total = p.m_blocks * p.n_blocks * p.split_k;
  const int sms = sm_count();
  if (sms <= 0) return FORDE_ENODEV;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(num_threads(EPI, PAIR));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  GemmParams q = p;
  q.clc = 0;
  if (PAIR) {
    // dynamic scheduling: one cluster per work item, resident clusters take over the not yet launched ones
    // (forde_gemm.cuh, "Tile scheduling"); FORDE_CLC=0 restores the static persistent grid for A/B runs
    static const bool clc_on = [] {
      const char* e = std::getenv("FORDE_CLC");
      return e == nullptr || std::atoi(e) != 0;
    }();
    int pairs = sms / 2;
    if (total < pairs) pairs = total;
    if (clc_on && total > pairs) {
      pairs = total;
      q.clc = 1;
    }
    cfg.gridDim = dim3(2 * pairs);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(total < sms ? total : sms);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, forde_gemm_kernel<A_MN, B_MN, EPI, PAIR>, q);
  if (e != cudaSuccess) {
    record_cuda_error("cudaLaunchKernelEx(forde_gemm_kernel)", cudaGetErrorString(e));
    return FORDE_ECUDA;
  }
  return check_launch("forde_gemm_kernel");
}

template <int EPI, bool PAIR>
static int launch_layouts(const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t stream) {
  if (a_mn) {
    if (b_mn) return launch_inst<true, true, EPI, PAIR>(p, stream);
    return launch_inst<true, false, EPI, PAIR>(p, stream);
  }
  if (b_mn) return launch_inst<false, true, EPI, PAIR>(p, stream);
  return launch_inst<false, false, EPI, PAIR>(p, stream);
}

template <int EPI>
static int launch_epi(const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t stream) {
  if (p.pair) return launch_layouts<EPI, true>(p, a_mn, b_mn, stream);
  return launch_layouts<EPI, false>(p, a_mn, b_mn, stream);
}

// CTA pairs (256-row tiles) once there are enough of them to fill the 74 SM pairs; single CTAs below that.
// FORDE_PAIR=0/1 overrides (bring-up and A/B measurements).
static bool choose_pair(int M, int N, int K, bool can_split_k) {
  static const int forced = [] {
    const char* e = std::getenv("FORDE_PAIR");
    return e ? std::atoi(e) : -1;
  }();
  if (forced >= 0) return forced != 0;
  long long work = static_cast<long long>((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  if (can_split_k) {  // split-K (fp32 accumulate outputs) multiplies the work items; >= 8 k-blocks per split
    const int max_split = ((K + BK - 1) / BK) / 8;
    if (max_split > 1) work *= max_split;
  }
  return work >= sm_count() / 2;
}

static int fill_common(GemmParams& p, const GemmArgs& g, bool can_split_k = false) {
  if (g.A == nullptr || g.B == nullptr || g.M <= 0 || g.N <= 0 || g.K <= 0) return FORDE_EINVAL;
  std::memset(&p, 0, sizeof(p));
  p.pair = choose_pair(g.M, g.N, g.K, can_split_k) ? 1 : 0;
  int rc = make_operand_map(&p.tmA, g.A, g.a_mn, g.lda, g.M, g.K, BM);
  if (rc != FORDE_OK) return rc;
  rc = make_operand_map(&p.tmB, g.B, g.b_mn, g.ldb, g.N, g.K, p.pair ? BN / 2 : BN);  // a pair CTA stages half of B
  if (rc != FORDE_OK) return rc;
  p.M = g.M; p.N = g.N; p.K = g.K;
  const int tile_m = p.pair ? 2 * BM : BM;
  p.m_blocks = (g.M + tile_m - 1) / tile_m;
  p.n_blocks = (g.N + BN - 1) / BN;
  p.k_blocks = (g.K + BK - 1) / BK;
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  p.raster_m_fast = g.raster_m_fast ? 1 : 0;
  p.hintA = g.hintA; p.hintB = g.hintB;
  p.scale = 1.f;
  p.gate_generalist = 1.0f;
  p.gate_specialist = 0.1f;
  return FORDE_OK;
}

static void choose_split_k(GemmParams& p) {
  const int tiles = p.m_blocks * p.n_blocks;
  const int workers = p.pair ? sm_count() / 2 : sm_count();
  int split = workers / tiles;
  const int max_split = p.k_blocks / 8 > 0 ? p.k_blocks / 8 : 1;  // at least 8 k-blocks per split
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  p.kb_per_split = (p.k_blocks + split - 1) / split;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
}

int gemm_store_bf16(cudaStream_t stream, const GemmArgs& g, void* D, long long ldd, const float* bias_col,
                    const void* Cin) {
  GemmParams p;
  int rc = fill_common(p, g);
  if (rc != FORDE_OK) return rc;
  if (D == nullptr) return FORDE_EINVAL;
  rc = make_tmap_bf16(&p.tmC0, D, g.N, g.M, ldd, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  p.bias_col = bias_col;
  if (Cin != nullptr) {  // D = A . B^T + bias + Cin (Cin has D's shape and pitch and may be D itself)
    rc = make_tmap_bf16(&p.tmZ, Cin, g.N, g.M, ldd, CHUNK, 32);
    if (rc != FORDE_OK) return rc;
    return launch_epi<EPI_STORE_ADD>(p, g.a_mn, g.b_mn, stream);
  }
  return launch_epi<EPI_STORE>(p, g.a_mn, g.b_mn, stream);
}

int gemm_f32(cudaStream_t stream, const GemmArgs& g, float* D, long long ldd, const float* row_scale,
             const float* col_scale, float scale, const float* scale_dev, bool accumulate) {
  GemmParams p;
  int rc = fill_common(p, g, accumulate);
  if (rc != FORDE_OK) return rc;
  if (D == nullptr) return FORDE_EINVAL;
  p.out_f32 = D; p.ldc = ldd; p.row_scale = row_scale; p.col_scale = col_scale; p.scale = scale; p.scale_dev = scale_dev;
  p.atomic_out = accumulate ? 1 : 0;
  if (accumulate) choose_split_k(p);
  return launch_epi<EPI_F32>(p, g.a_mn, g.b_mn, stream);
}

// ------------------------------------------------------------------ grouped launches (MoE experts)
template <bool A_MN, bool B_MN, int EPI>
static int launch_single(const GemmParams& p, cudaStream_t stream) {
  return launch_inst<A_MN, B_MN, EPI, false>(p, stream);
}

int gemm_grouped_m(cudaStream_t stream, const GemmArgs& g, const GroupedM& gm, int mode, void* D, long long ldd,
                   void* D2, const void* Zin, const float* bias_col) {
  if (g.a_mn || D == nullptr || gm.m_blocks_dev == nullptr || gm.group_of_mblock == nullptr) return FORDE_EINVAL;
  GemmParams p;
  int rc = fill_common(p, g);
  if (rc != FORDE_OK) return rc;
  // grouped kernels are single-CTA instantiations (128-row tiles), and B is the expert-STACKED weight array:
  // rebuild its map with the stacked extents
  p.pair = 0;
  p.m_blocks = (g.M + BM - 1) / BM;
  {
    const int rows = g.N * (gm.b_n_group_stride ? gm.groups : 1);
    const int kext = g.K * (gm.b_k_group_stride ? gm.groups : 1);
    rc = make_operand_map(&p.tmB, g.B, g.b_mn, g.ldb, rows, kext, BN);
    if (rc != FORDE_OK) return rc;
  }
  rc = make_tmap_bf16(&p.tmC0, D, g.N, g.M, ldd, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  p.bias_col = bias_col;
  p.m_blocks_dev = gm.m_blocks_dev;
  p.group_of_mblock = gm.group_of_mblock;
  p.b_k_group_stride = gm.b_k_group_stride;
  p.b_n_group_stride = gm.b_n_group_stride;
  p.bias_group_stride = gm.bias_group_stride;
  p.raster_m_fast = 0;
  if (mode == 1) {
    if (D2 == nullptr) return FORDE_EINVAL;
    rc = make_tmap_bf16(&p.tmC1, D2, g.N, g.M, ldd, CHUNK, 32);
    if (rc != FORDE_OK) return rc;
    return g.b_mn ? launch_single<false, true, EPI_MOE_UP>(p, stream) : launch_single<false, false, EPI_MOE_UP>(p, stream);
  }
  if (mode == 2) {
    if (Zin == nullptr) return FORDE_EINVAL;
    rc = make_tmap_bf16(&p.tmZ, Zin, g.N, g.M, ldd, CHUNK, 32);
    if (rc != FORDE_OK) return rc;
    return g.b_mn ? launch_single<false, true, EPI_MOE_DZ>(p, stream) : launch_single<false, false, EPI_MOE_DZ>(p, stream);
  }
  return g.b_mn ? launch_single<false, true, EPI_STORE>(p, stream) : launch_single<false, false, EPI_STORE>(p, stream);
}

int gemm_grouped_k(cudaStream_t stream, const GemmArgs& g, const int* seg_offsets, int groups, float* out,
                   long long ldo, long long out_group_stride) {
  if (!g.a_mn || !g.b_mn || out == nullptr || seg_offsets == nullptr || groups <= 0) return FORDE_EINVAL;
  GemmParams p;
  int rc = fill_common(p, g);
  if (rc != FORDE_OK) return rc;
  if (p.pair) {
    p.pair = 0;
    rc = make_operand_map(&p.tmB, g.B, g.b_mn, g.ldb, g.N, g.K, BN);
    if (rc != FORDE_OK) return rc;
    p.m_blocks = (g.M + BM - 1) / BM;
  }
  p.out_f32 = out; p.ldc = ldo; p.scale = 1.f; p.atomic_out = 0;
  p.seg_offsets = seg_offsets; p.groups = groups; p.out_group_stride = out_group_stride;
  // total work = groups * tiles: tell the launcher through split_k so the grid is sized for it
  p.split_k = groups;
  return launch_single<true, true, EPI_F32>(p, stream);
}

}  // namespace forde

using namespace forde;

extern "C" {

int forde_version(void) { return 100; }

const char* forde_error_string(int code) {
  switch (code) {
    case FORDE_OK: return "ok";
    case FORDE_EINVAL: return "invalid argument (null pointer or non-positive dimension)";
    case FORDE_EALIGN: return "alignment violation (bf16 matrices need 16-byte bases and pitches that are multiples of 8)";
    case FORDE_ECUDA: return "CUDA error (see forde_last_cuda_error)";
    case FORDE_ENODEV: return "no CUDA device";
    default: return "unknown error";
  }
}

const char* forde_last_cuda_error(void) { return g_last_err; }

int forde_sm_count(void) { return sm_count(); }

unsigned long long forde_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int forde_sl_sign_layout(void) { return 3; }  // bit 0: interleaved bit order ; bit 1: word-major buffer

int forde_sl_sign_words(int M) { return M <= 0 ? 0 : 8 * ((M + BN - 1) / BN); }

int forde_sl_num_partials(int M) { return M <= 0 ? 0 : PARTIALS_PER_TILE * ((M + BN - 1) / BN); }

int forde_gemm_bf16(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                    long long ldb, void* D, long long ldd, const float* bias_col, int M, int N, int K) {
  GemmArgs g{A, a_mn_major != 0, lda, B, b_mn_major != 0, ldb, M, N, K, false, kEvictNormal, kEvictNormal};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, D, ldd, bias_col);
}

int forde_gemm_bf16_add(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                        long long ldb, void* D, long long ldd, const void* Cin, const float* bias_col, int M, int N,
                        int K) {
  if (Cin == nullptr) return FORDE_EINVAL;
  GemmArgs g{A, a_mn_major != 0, lda, B, b_mn_major != 0, ldb, M, N, K, false, kEvictNormal, kEvictNormal};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, D, ldd, bias_col, Cin);
}

int forde_gemm_f32(void* stream, const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                   long long ldb, float* D, long long ldd, const float* row_scale, const float* col_scale,
                   float scale, int accumulate, int M, int N, int K) {
  GemmArgs g{A, a_mn_major != 0, lda, B, b_mn_major != 0, ldb, M, N, K, false, kEvictNormal, kEvictNormal};
  return gemm_f32(static_cast<cudaStream_t>(stream), g, D, ldd, row_scale, col_scale, scale, nullptr, accumulate != 0);
}

// zT[H][M] = (x . w_up + b)^T :  A = w_up [d][H] as (m = neuron, k = d) -> MN-major ; B = x [M][d] -> K-major
static int sl_up_fwd_impl(void* stream, const void* x, const void* w_up, const float* b_up, const int8_t* assign,
                          void* zT, void* hT, uint32_t* sign_bits, float* act_partials, int M, int d, int H,
                          long long ldz, int save_z) {
  if (assign == nullptr || zT == nullptr || hT == nullptr || ldz < M) return FORDE_EINVAL;
  if (sign_bits != nullptr && (reinterpret_cast<uintptr_t>(sign_bits) & 15) != 0) return FORDE_EALIGN;
  GemmArgs g{w_up, true, H, x, false, d, H, M, d, /*raster_m_fast=*/true, kEvictLast, kEvictLast};
  GemmParams p;
  int rc = fill_common(p, g);
  if (rc != FORDE_OK) return rc;
  rc = make_tmap_bf16(&p.tmC0, zT, M, H, ldz, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  rc = make_tmap_bf16(&p.tmC1, hT, M, H, ldz, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  p.bias_row = b_up;
  p.assign = assign;
  p.partials = act_partials;
  p.save_z = save_z;
  p.sign_bits = sign_bits;
  p.sign_ld = forde_sl_sign_words(M);
  return launch_epi<EPI_UP_FWD>(p, true, false, static_cast<cudaStream_t>(stream));
}

int forde_sl_up_fwd(void* stream, const void* x, const void* w_up, const float* b_up, const int8_t* assign,
                    void* zT, void* hT, float* act_partials, int M, int d, int H, long long ldz, int save_z) {
  return sl_up_fwd_impl(stream, x, w_up, b_up, assign, zT, hT, nullptr, act_partials, M, d, H, ldz, save_z);
}

int forde_sl_up_fwd_signs(void* stream, const void* x, const void* w_up, const float* b_up, const int8_t* assign,
                          void* zT, void* hT, uint32_t* sign_bits, float* act_partials, int M, int d, int H,
                          long long ldz) {
  if (sign_bits == nullptr) return FORDE_EINVAL;
  return sl_up_fwd_impl(stream, x, w_up, b_up, assign, zT, hT, sign_bits, act_partials, M, d, H, ldz, /*save_z=*/0);
}

// y[M][d] = h . w_down + b :  A = hT [H][M] as (m = sample, k = neuron) -> MN-major ; B = w_down [H][d] -> MN-major
int forde_sl_down_fwd(void* stream, const void* hT, const void* w_down, const float* b_down, void* y, int M,
                      int d, int H, long long ldz) {
  if (ldz < M) return FORDE_EINVAL;
  GemmArgs g{hT, true, ldz, w_down, true, d, M, d, H, /*raster_m_fast=*/false, kEvictFirst, kEvictLast};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, y, d, b_down);
}

// y = h . w_down + b + residual : the block's `x + StatefulLayer(x)` is formed in the epilogue, so the add/norm that
// follows reads ONE tensor instead of two
int forde_sl_down_fwd_add(void* stream, const void* hT, const void* w_down, const float* b_down, const void* residual,
                          void* y, int M, int d, int H, long long ldz) {
  if (ldz < M || residual == nullptr) return FORDE_EINVAL;
  GemmArgs g{hT, true, ldz, w_down, true, d, M, d, H, /*raster_m_fast=*/false, kEvictFirst, kEvictLast};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, y, d, b_down, residual);
}

int forde_sl_fwd(void* stream, const void* x, const void* w_up, const float* b_up, const void* w_down,
                 const float* b_down, const int8_t* assign, void* zT, void* hT, void* y, float* act_partials,
                 int M, int d, int H, long long ldz, int save_z) {
  int rc = forde_sl_up_fwd(stream, x, w_up, b_up, assign, zT, hT, act_partials, M, d, H, ldz, save_z);
  if (rc != FORDE_OK) return rc;
  return forde_sl_down_fwd(stream, hT, w_down, b_down, y, M, d, H, ldz);
}

// dzT[H][M] = (dy . w_down^T)^T * (F'(z) + gate) : A = w_down [H][d] (m = neuron, k = d) K-major ; B = dy [M][d] K-major
static int sl_dz_bwd_impl(void* stream, const void* dy, const void* w_down, const void* zT, const void* hT,
                          const uint32_t* sign_bits, const int8_t* assign, void* dzT, float* grad_partials, int M,
                          int d, int H, long long ldz, const uint32_t* keep_bits = nullptr) {
  if (assign == nullptr || zT == nullptr || dzT == nullptr || ldz < M) return FORDE_EINVAL;
  if (hT == nullptr && sign_bits == nullptr) return FORDE_EINVAL;
  if (sign_bits != nullptr && (reinterpret_cast<uintptr_t>(sign_bits) & 15) != 0) return FORDE_EALIGN;
  if (hT == nullptr) hT = zT;  // sign-bit mode never loads through this map; it only has to encode
  GemmArgs g{w_down, false, d, dy, false, d, H, M, d, /*raster_m_fast=*/true, kEvictLast, kEvictLast};
  GemmParams p;
  int rc = fill_common(p, g);
  if (rc != FORDE_OK) return rc;
  rc = make_tmap_bf16(&p.tmC0, dzT, M, H, ldz, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  rc = make_tmap_bf16(&p.tmZ, zT, M, H, ldz, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  rc = make_tmap_bf16(&p.tmH, hT, M, H, ldz, CHUNK, 32);
  if (rc != FORDE_OK) return rc;
  p.assign = assign;
  p.partials = grad_partials;
  p.sign_bits = const_cast<uint32_t*>(sign_bits);
  p.sign_ld = forde_sl_sign_words(M);
  p.keep_bits = keep_bits;
  return launch_epi<EPI_DZ_BWD>(p, false, false, static_cast<cudaStream_t>(stream));
}

int forde_sl_dz_bwd(void* stream, const void* dy, const void* w_down, const void* zT, const void* hT,
                    const int8_t* assign, void* dzT, float* grad_partials, int M, int d, int H, long long ldz) {
  if (hT == nullptr) return FORDE_EINVAL;
  return sl_dz_bwd_impl(stream, dy, w_down, zT, hT, nullptr, assign, dzT, grad_partials, M, d, H, ldz);
}

int forde_sl_dz_bwd_signs(void* stream, const void* dy, const void* w_down, const void* zT,
                          const uint32_t* sign_bits, const int8_t* assign, void* dzT, float* grad_partials, int M,
                          int d, int H, long long ldz) {
  if (sign_bits == nullptr) return FORDE_EINVAL;
  return sl_dz_bwd_impl(stream, dy, w_down, zT, nullptr, sign_bits, assign, dzT, grad_partials, M, d, H, ldz);
}

int forde_sl_dz_bwd_topk(void* stream, const void* dy, const void* w_down, const void* zT, const void* hT,
                         const uint32_t* sign_bits, const uint32_t* keep_bits, const int8_t* assign, void* dzT,
                         float* grad_partials, int M, int d, int H, long long ldz) {
  if (keep_bits == nullptr || (hT == nullptr && sign_bits == nullptr)) return FORDE_EINVAL;
  return sl_dz_bwd_impl(stream, dy, w_down, zT, sign_bits != nullptr ? nullptr : hT, sign_bits, assign, dzT,
                        grad_partials, M, d, H, ldz, keep_bits);
}

// dx[M][d] = dz . w_up^T : A = dzT [H][M] (m = sample, k = neuron) MN-major ; B = w_up [d][H] (n = d, k = neuron) K-major
int forde_sl_dx_bwd(void* stream, const void* dzT, const void* w_up, void* dx, int M, int d, int H,
                    long long ldz) {
  if (ldz < M) return FORDE_EINVAL;
  GemmArgs g{dzT, true, ldz, w_up, false, H, M, d, H, /*raster_m_fast=*/false, kEvictFirst, kEvictLast};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, dx, d, nullptr);
}

// dx = dz . w_up^T + residual : the FORDE block's `x + StatefulLayer(x)` skip path joins its gradient here
int forde_sl_dx_bwd_add(void* stream, const void* dzT, const void* w_up, const void* residual, void* dx, int M, int d,
                        int H, long long ldz) {
  if (ldz < M || residual == nullptr) return FORDE_EINVAL;
  GemmArgs g{dzT, true, ldz, w_up, false, H, M, d, H, /*raster_m_fast=*/false, kEvictFirst, kEvictLast};
  return gemm_store_bf16(static_cast<cudaStream_t>(stream), g, dx, d, nullptr, residual);
}

// dw_up[d][H] += x^T . dz : A = x [M][d] (m = d, k = sample) MN-major ; B = dzT [H][M] (n = neuron, k = sample) K-major
int forde_sl_dw_up(void* stream, const void* x, const void* dzT, float* dw_up, int M, int d, int H,
                   long long ldz) {
  if (ldz < M) return FORDE_EINVAL;
  GemmArgs g{x, true, d, dzT, false, ldz, d, H, M, /*raster_m_fast=*/true, kEvictNormal, kEvictNormal};
  return gemm_f32(static_cast<cudaStream_t>(stream), g, dw_up, H, nullptr, nullptr, 1.f, nullptr, true);
}

// dw_down[H][d] += h^T . dy : A = hT [H][M] (m = neuron, k = sample) K-major ; B = dy [M][d] (n = d, k = sample) MN-major
int forde_sl_dw_down(void* stream, const void* hT, const void* dy, float* dw_down, int M, int d, int H,
                     long long ldz) {
  if (ldz < M) return FORDE_EINVAL;
  GemmArgs g{hT, false, ldz, dy, true, d, H, d, M, /*raster_m_fast=*/false, kEvictNormal, kEvictNormal};
  return gemm_f32(static_cast<cudaStream_t>(stream), g, dw_down, d, nullptr, nullptr, 1.f, nullptr, true);
}

int forde_sl_bwd(void* stream, const void* x, const void* dy, const void* w_up, const void* w_down,
                 const void* zT, const void* hT, const int8_t* assign, void* dzT, void* dx, float* dw_up,
                 float* db_up, float* dw_down, float* db_down, float* grad_partials, int M, int d, int H,
                 long long ldz) {
  int rc = forde_sl_dz_bwd(stream, dy, w_down, zT, hT, assign, dzT, grad_partials, M, d, H, ldz);
  if (rc != FORDE_OK) return rc;
  if (dx != nullptr) {
    rc = forde_sl_dx_bwd(stream, dzT, w_up, dx, M, d, H, ldz);
    if (rc != FORDE_OK) return rc;
  }
  if (dw_up != nullptr) {
    rc = forde_sl_dw_up(stream, x, dzT, dw_up, M, d, H, ldz);
    if (rc != FORDE_OK) return rc;
  }
  if (dw_down != nullptr) {
    rc = forde_sl_dw_down(stream, hT, dy, dw_down, M, d, H, ldz);
    if (rc != FORDE_OK) return rc;
  }
  if (db_down != nullptr) {
    rc = forde_colsum_bf16(stream, dy, d, db_down, M, d);
    if (rc != FORDE_OK) return rc;
  }
  if (db_up != nullptr && grad_partials != nullptr) {
    rc = forde_stats_finalize(stream, nullptr, grad_partials, forde_sl_num_partials(M), H, M, nullptr, nullptr,
                              nullptr, nullptr, nullptr, nullptr, db_up);
    if (rc != FORDE_OK) return rc;
  }
  return FORDE_OK;
}

}  // extern "C"
