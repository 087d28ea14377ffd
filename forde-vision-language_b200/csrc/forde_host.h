// Host-side helpers shared between the translation units of libforde_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace forde {

void record_cuda_error(const char* what, const char* msg);
// records a launch error (if any) and counts `kernels` successful kernel launches (forde_launch_count)
int check_launch(const char* what, int kernels = 1);
int sm_count();

struct GemmArgs {
  const void* A;
  bool a_mn;       // A stored [K][M] (true) or [M][K] (false)
  long long lda;
  const void* B;
  bool b_mn;       // B stored [K][N] (true) or [N][K] (false)
  long long ldb;
  int M, N, K;
  bool raster_m_fast;
  uint64_t hintA, hintB;
};

// D (bf16) = A . B^T (+ bias_col)
int gemm_store_bf16(cudaStream_t stream, const GemmArgs& g, void* D, long long ldd, const float* bias_col);
// D (fp32) = or += scale * row_scale[m] * col_scale[n] * (A . B^T); accumulate uses split-K + red.add
int gemm_f32(cudaStream_t stream, const GemmArgs& g, float* D, long long ldd, const float* row_scale,
             const float* col_scale, float scale, const float* scale_dev, bool accumulate);

}  // namespace forde
