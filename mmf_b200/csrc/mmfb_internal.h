// Internal host-side declarations shared by the translation units of libmmfb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmfb200.h"

namespace mmfb {

enum {
  EPI_BIAS = MMFB_EPI_BIAS,
  EPI_BIAS_GELU = MMFB_EPI_BIAS_GELU,
  EPI_BIAS_DROP_RESID = MMFB_EPI_BIAS_DROP_RESID,
  EPI_GELU_BWD = MMFB_EPI_GELU_BWD,
  EPI_ADD_AUX = MMFB_EPI_ADD_AUX,
  EPI_ATOMIC_F32 = MMFB_EPI_ATOMIC_F32,
  EPI_BIAS_RELU = MMFB_EPI_BIAS_RELU
};

// records a thread-local error message and returns `code`
int set_error(int code, const char* fmt, ...);
int num_sms();
void count_launch(int n = 1);

// bf16 2-D tensor map: tensor [outer, inner] with row stride `ld` elements, box {box_inner, box_outer},
// SWIZZLE_128B (box_inner must be 64), out-of-bounds elements read as zero / are not written.
int make_tmap_2d(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                 int box_outer);
// bf16 3-D tensor map: tensor [d2, d1, inner] with strides ld1 (rows) and ld2 (batches) in elements
int make_tmap_3d(CUtensorMap* map, const void* ptr, int64_t inner, int64_t d1, int64_t d2, int64_t ld1,
                 int64_t ld2, int box_inner, int box_d1);

// Launch with the programmatic-stream-serialization attribute (common.cuh: griddep_launch / griddep_wait), optionally as
// clusters of `cluster` CTAs.  MMFB_PDL=0 (read once) launches plainly - the A/B switch.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define MMFB_LAUNCH(kern, grid, block, smem, stream, ...) \
  (void)::mmfb::launch_k(kern, dim3(grid), dim3(block), static_cast<size_t>(smem), stream, 1, __VA_ARGS__)

int gemm(const mmfb_gemm_args& a, cudaStream_t stream);
int attn_fwd(const mmfb_attn_args& a, cudaStream_t stream);
int attn_bwd(const mmfb_attn_args& a, cudaStream_t stream);
int ln_fwd(const mmfb_ln_args& a, cudaStream_t s);
int ln_bwd(const mmfb_ln_args& a, cudaStream_t s);
int colsum(const void* X, int64_t ldx, float* out, int M, int N, cudaStream_t s);
int dropout_bits(uint32_t* out, int64_t nwords, uint64_t seed, uint64_t offset, float p, const uint64_t* epoch, cudaStream_t s);
int compose(const mmfb_compose_args& a, cudaStream_t s);
int scatter(const mmfb_scatter_args& a, cudaStream_t s);
int cast_params(const float* in, void* out, int64_t n, cudaStream_t s);
int relu_bwd(const void* dy, const void* y, void* dz, int64_t n, cudaStream_t s);
int adamw(const mmfb_adamw_args& a, cudaStream_t s);
int ce_rows(void* logits, int64_t ldl, const int64_t* labels, int64_t ignore_index, int M, int V, float grad_scale,
            float* loss_sum, float* row_loss, cudaStream_t s);
int add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t s);
int dropout_apply(const void* x, int64_t ldx, const uint32_t* bits, int64_t ldm, float scale, void* out, int64_t ldo, int M,
                  int H, cudaStream_t s);
int gelu_bwd(const void* dh, const void* u, void* du, int64_t n, cudaStream_t s);
int scatter_sorted(const void* dy, int64_t lddy, const int32_t* order, const int32_t* sorted_idx, float* dtab, int M,
                   int H, cudaStream_t s);

}  // namespace mmfb
