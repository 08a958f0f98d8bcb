// Fused AdamW over the flat fp32 parameter / gradient buffers of the engine (include/mmfb200.h: mmfb_adamw).
// HBM-bound: 4 fp32 streams read + 3 written (+ the bf16 compute copy) = 30 bytes per parameter, one pass.
// Arithmetic follows the reference's optimizer exactly, operation by operation (mmf/modules/optimizers.py:60-84 for the
// transformers variant, torch/optim/adamw.py single-tensor path for the fallback), so that parity is an fp32 statement.
#include "common.cuh"
#include "mmfb_internal.h"

namespace mmfb {

struct AdamWDev {
  float* p; const float* g; float* m; float* v; bf16* pb; const uint8_t* grp; int64_t n;
  float lr[MMFB_ADAMW_MAX_GROUPS], wd[MMFB_ADAMW_MAX_GROUPS], step[MMFB_ADAMW_MAX_GROUPS], bc2s[MMFB_ADAMW_MAX_GROUPS];
  float b1, b2, eps, gscale;
};

template <int MODE>
__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ AdamWDev a) {
  griddep_launch();
  griddep_wait();
  const int64_t blk = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;     // 8-element block
  const int64_t i = blk * 8;
  if (i >= a.n) return;
  const int gi = a.grp != nullptr ? a.grp[blk] : 0;
  const float lr = a.lr[gi], wd = a.wd[gi], step = a.step[gi], bc2s = a.bc2s[gi];
  float p[8], g[8], m[8], v[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 tp = *reinterpret_cast<const float4*>(a.p + i + 4 * h);
    const float4 tg = *reinterpret_cast<const float4*>(a.g + i + 4 * h);
    const float4 tm = *reinterpret_cast<const float4*>(a.m + i + 4 * h);
    const float4 tv = *reinterpret_cast<const float4*>(a.v + i + 4 * h);
    p[4 * h] = tp.x; p[4 * h + 1] = tp.y; p[4 * h + 2] = tp.z; p[4 * h + 3] = tp.w;
    g[4 * h] = tg.x; g[4 * h + 1] = tg.y; g[4 * h + 2] = tg.z; g[4 * h + 3] = tg.w;
    m[4 * h] = tm.x; m[4 * h + 1] = tm.y; m[4 * h + 2] = tm.z; m[4 * h + 3] = tm.w;
    v[4 * h] = tv.x; v[4 * h + 1] = tv.y; v[4 * h + 2] = tv.z; v[4 * h + 3] = tv.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gr = g[e] * a.gscale;
    if (MODE == 0) {
      // exp_avg.mul_(b1).add_(g, alpha=1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
      m[e] = __fmaf_rn(gr, 1.0f - a.b1, m[e] * a.b1);
      v[e] = __fmaf_rn(gr * gr, 1.0f - a.b2, v[e] * a.b2);
      const float denom = __fsqrt_rn(v[e]) + a.eps;
      p[e] = __fmaf_rn(-step, __fdiv_rn(m[e], denom), p[e]);              // p.addcdiv_(m, denom, value=-step_size)
      if (wd > 0.0f) p[e] = __fmaf_rn(p[e], -lr * wd, p[e]);               // p.add_(p, alpha=-lr*wd)
    } else {
      p[e] = p[e] * (1.0f - lr * wd);                                      // p.mul_(1 - lr*wd)
      m[e] = __fmaf_rn(gr - m[e], 1.0f - a.b1, m[e]);                      // m.lerp_(g, 1-b1)
      v[e] = __fmaf_rn(gr * gr, 1.0f - a.b2, v[e] * a.b2);
      const float denom = __fdiv_rn(__fsqrt_rn(v[e]), bc2s) + a.eps;
      p[e] = __fmaf_rn(-step, __fdiv_rn(m[e], denom), p[e]);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    *reinterpret_cast<float4*>(a.p + i + 4 * h) = make_float4(p[4 * h], p[4 * h + 1], p[4 * h + 2], p[4 * h + 3]);
    *reinterpret_cast<float4*>(a.m + i + 4 * h) = make_float4(m[4 * h], m[4 * h + 1], m[4 * h + 2], m[4 * h + 3]);
    *reinterpret_cast<float4*>(a.v + i + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
  }
  if (a.pb != nullptr) {
    uint4 o;
    o.x = pack_bf16x2(p[0], p[1]); o.y = pack_bf16x2(p[2], p[3]);
    o.z = pack_bf16x2(p[4], p[5]); o.w = pack_bf16x2(p[6], p[7]);
    *reinterpret_cast<uint4*>(a.pb + i) = o;
  }
}

int adamw(const mmfb_adamw_args& a, cudaStream_t s) {
  if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq) return set_error(MMFB_ERR_ARG, "adamw: null pointer");
  if (a.n <= 0 || (a.n % 8)) return set_error(MMFB_ERR_ARG, "adamw: n must be a positive multiple of 8 (got %lld)", (long long)a.n);
  if (a.n_groups < 1 || a.n_groups > MMFB_ADAMW_MAX_GROUPS)
    return set_error(MMFB_ERR_ARG, "adamw: n_groups must be in [1, %d]", MMFB_ADAMW_MAX_GROUPS);
  if (a.mode != 0 && a.mode != 1) return set_error(MMFB_ERR_ARG, "adamw: mode must be 0 (transformers) or 1 (torch)");
  if ((reinterpret_cast<uintptr_t>(a.param) | reinterpret_cast<uintptr_t>(a.grad) | reinterpret_cast<uintptr_t>(a.exp_avg) |
       reinterpret_cast<uintptr_t>(a.exp_avg_sq) | reinterpret_cast<uintptr_t>(a.param_bf16)) & 15)
    return set_error(MMFB_ERR_ARG, "adamw: buffers must be 16-byte aligned");
  AdamWDev d;
  d.p = a.param; d.g = a.grad; d.m = a.exp_avg; d.v = a.exp_avg_sq; d.pb = reinterpret_cast<bf16*>(a.param_bf16);
  d.grp = a.group; d.n = a.n;
  for (int i = 0; i < MMFB_ADAMW_MAX_GROUPS; ++i) {
    const int j = i < a.n_groups ? i : 0;
    d.lr[i] = a.lr[j]; d.wd[i] = a.weight_decay[j]; d.step[i] = a.step_size[j]; d.bc2s[i] = a.bc2_sqrt[j];
  }
  d.b1 = a.beta1; d.b2 = a.beta2; d.eps = a.eps; d.gscale = a.grad_scale;
  const int64_t blocks8 = a.n / 8;
  const unsigned grid = static_cast<unsigned>((blocks8 + 255) / 256);
  if (a.mode == 0) MMFB_LAUNCH(adamw_kernel<0>, grid, 256, 0, s, d);
  else MMFB_LAUNCH(adamw_kernel<1>, grid, 256, 0, s, d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MMFB_ERR_CUDA, "adamw launch: %s", cudaGetErrorString(e));
  count_launch();
  return MMFB_OK;
}

// du = dh * GELU'(u), 16-byte vectors (staged with the MLM head)
__global__ void gelu_bwd_kernel(const bf16* __restrict__ dh, const bf16* __restrict__ u, bf16* __restrict__ du, int64_t n) {
  griddep_launch();
  griddep_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 a = *reinterpret_cast<const uint4*>(dh + i);
    const uint4 b = *reinterpret_cast<const uint4*>(u + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 d = unpack_bf16x2(aw[k]), x = unpack_bf16x2(bw[k]);
      ow[k] = pack_bf16x2(d.x * gelu_erf_grad(x.x), d.y * gelu_erf_grad(x.y));
    }
    *reinterpret_cast<uint4*>(du + i) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  } else {
    for (int64_t j = i; j < n; ++j)
      du[j] = __float2bfloat16_rn(__bfloat162float(dh[j]) * gelu_erf_grad(__bfloat162float(u[j])));
  }
}

int gelu_bwd(const void* dh, const void* u, void* du, int64_t n, cudaStream_t s) {
  if (n <= 0) return set_error(MMFB_ERR_ARG, "gelu_bwd: empty");
  if ((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(du)) & 15)
    return set_error(MMFB_ERR_ARG, "gelu_bwd: buffers must be 16-byte aligned");
  const int64_t thr = (n + 7) / 8;
  MMFB_LAUNCH(gelu_bwd_kernel, static_cast<unsigned>((thr + 255) / 256), 256, 0, s, (const bf16*)dh, (const bf16*)u, (bf16*)du, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MMFB_ERR_CUDA, "gelu_bwd launch: %s", cudaGetErrorString(e));
  count_launch();
  return MMFB_OK;
}

}  // namespace mmfb
