// Host-side helpers shared between the translation units of libforde_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace forde {

void record_cuda_error(const char* what, const char* msg);
// records a launch error (if any) and counts `kernels` successful kernel launches (forde_launch_count)
int check_launch(const char* what, int kernels = 1);
int sm_count();

// cached TMA descriptor of a 3-D bf16 tensor (see forde_api.cu)
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                      uint64_t ld2, uint32_t b0, uint32_t b1);

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

// D (bf16) = A . B^T (+ bias_col) (+ Cin: same shape / pitch as D, may alias it)
int gemm_store_bf16(cudaStream_t stream, const GemmArgs& g, void* D, long long ldd, const float* bias_col,
                    const void* Cin = nullptr);
// D (fp32) = or += scale * row_scale[m] * col_scale[n] * (A . B^T); accumulate uses split-K + red.add
int gemm_f32(cudaStream_t stream, const GemmArgs& g, float* D, long long ldd, const float* row_scale,
             const float* col_scale, float scale, const float* scale_dev, bool accumulate);

// ---- grouped GEMMs for the MoE experts (rows = routed token slots sorted by expert, segments padded to 128) -------
struct GroupedM {
  const int* m_blocks_dev;     // device scalar: number of valid 128-row blocks
  const int* group_of_mblock;  // expert of each 128-row block
  int b_k_group_stride;        // B stacked along k: k offset per expert (else 0)
  int b_n_group_stride;        // B stacked along n: row offset per expert (else 0)
  int bias_group_stride;       // bias_col offset per expert
  int groups;                  // number of experts stacked in B
};
// D (bf16, [M_cap][N]) = A . B_e^T + bias_e ; optional second output / z input select the MoE epilogues:
//   mode 0: plain store ; mode 1: EPI_MOE_UP (D = z, D2 = gelu(z)) ; mode 2: EPI_MOE_DZ (D = acc * gelu'(Zin))
int gemm_grouped_m(cudaStream_t stream, const GemmArgs& g, const GroupedM& gm, int mode, void* D, long long ldd,
                   void* D2, const void* Zin, const float* bias_col);
// out[e] (fp32 [M][N], plain store) = sum over expert e's row segment of A(m, row) * B(n, row); A and B MN-major.
int gemm_grouped_k(cudaStream_t stream, const GemmArgs& g, const int* seg_offsets, int groups, float* out,
                   long long ldo, long long out_group_stride);

}  // namespace forde
