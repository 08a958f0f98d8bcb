// Row-wise (HBM-bound) kernels of the fusion block: LayerNorm forward/backward (with the dropout
// backward and bias-gradient column sums folded in), bias-gradient column sums, Philox dropout keep-bit
// generation, the embedding "compose + LayerNorm" kernel (region-feature / token gather + type /
// position add + LN) and its scatter backward, fp32 -> bf16 parameter cast.
//
// All of them are one-warp-per-row, 16-byte vectorised, fully coalesced; nothing here is GEMM-shaped.
// Reference ops: nn.LayerNorm(eps=1e-12) in HF BertSelfOutput/BertOutput (called at
// mmf/modules/hf_layers.py:248,290) and mmf/models/vilbert.py:254,303,483,490,911; nn.Dropout at the same
// sites; embeddings mmf/modules/embeddings.py:329-370,423-459, mmf/models/mmbt.py:92-129,
// mmf/models/transformers/backends/huggingface.py:131-159, mmf/models/vilbert.py:904-913.
#include "common.cuh"
#include "mmfb_internal.h"

namespace mmfb {

constexpr int MAXV = 8;         // per-lane 8-element vectors: rows up to 8*256 = 2048 columns
constexpr int ROW_WARPS = 8;    // warps per block

__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm forward:  x = LN(y) * gamma + beta  (+ optional dropout on the OUTPUT, used by the embeddings)
// ----------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_fwd_kernel(const bf16* __restrict__ y, int64_t ldy, const bf16* __restrict__ gamma,
              const bf16* __restrict__ beta, bf16* __restrict__ x, int64_t ldx, float* __restrict__ mean,
              float* __restrict__ rstd, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale, int M,
              int H, float eps) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[NV][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      ld8(y + row * ldy + col, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mu = warp_sum(s) / H;
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mu; ss += d * d; }
    }
  }
  const float rs = rsqrtf(warp_sum(ss) / H + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      float g[8], b[8], o[8];
      ld8(gamma + col, g);
      ld8(beta + col, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mu) * rs * g[e] + b[e];
      if (dmask != nullptr) {
        const uint32_t w = __ldg(dmask + row * ldmask + (col >> 5));
        const uint32_t bits = (w >> (col & 31)) & 0xFFu;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ((bits >> e) & 1u) ? o[e] * dscale : 0.0f;
      }
      st8(x + row * ldx + col, o);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// LayerNorm backward (+ dropout backward of the pre-LN dense branch + bias gradient)
//   dy    = gamma*rstd * (dx - mean_h(dx*gamma) ... ) standard LN backward wrt the pre-LN sum y = z + resid
//   dz    = dy * keep * dscale        (gradient of the dense output that was dropped)   [== dy when no dropout]
//   dgamma += sum_rows dx * xhat ; dbeta += sum_rows dx ; dbias += sum_rows dz
// dx_in may carry an extra additive term dx2 (the residual gradient arriving from a later consumer).
// A persistent grid (grid-stride over row groups) keeps the number of fp32 atomics at 3*H per block.
// ----------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_bwd_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
              const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
              const float* __restrict__ rstd, const bf16* __restrict__ gamma, bf16* __restrict__ dy, int64_t lddy,
              bf16* __restrict__ dz, int64_t lddz, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int H) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[ROW_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[NV][8], ab[NV][8], az[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { ag[i][e] = 0.0f; ab[i][e] = 0.0f; az[i][e] = 0.0f; }

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + warp; row < M;
       row += static_cast<int64_t>(gridDim.x) * ROW_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    float g[NV][8], d[NV][8], xh[NV][8];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        float yy[8];
        ld8(dx + row * lddx + col, d[i]);
        if (dx2 != nullptr) {
          float t[8];
          ld8(dx2 + row * lddx2 + col, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[i][e] += t[e];
        }
        ld8(y + row * ldy + col, yy);
        ld8(gamma + col, g[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[i][e] = (yy[e] - mu) * rs;
          const float dg = d[i][e] * g[i][e];
          s1 += dg;
          s2 += dg * xh[i][e];
          ag[i][e] += d[i][e] * xh[i][e];
          ab[i][e] += d[i][e];
        }
      }
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (d[i][e] * g[i][e] - s1 - xh[i][e] * s2);
        if (dy != nullptr) st8(dy + row * lddy + col, o);
        if (dmask != nullptr) {
          const uint32_t w = __ldg(dmask + row * ldmask + (col >> 5));
          const uint32_t bits = (w >> (col & 31)) & 0xFFu;
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = ((bits >> e) & 1u) ? o[e] * dscale : 0.0f;
          st8(dz + row * lddz + col, o);
        } else if (dz != nullptr && dz != dy) {
          st8(dz + row * lddz + col, o);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) az[i][e] += o[e];
      }
    }
  }
  // block reduction of the three column-sum sets, 256 columns at a time, then one atomic per column
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 256 >= H) break;
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = which == 0 ? ag[i][e] : (which == 1 ? ab[i][e] : az[i][e]);
      __syncthreads();
      const int c = threadIdx.x;  // 256 threads <-> 256 columns
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) t += red[w][c];
      if (i * 256 + c < H) atomicAdd(dst + i * 256 + c, t);
    }
  }
}

// Single-pass LayerNorm backward, lean variant (staged: selected with MMFB_LN_BWD=lean, not the default until it is
// measured).  Same results as ln_bwd_kernel, built for two 256-thread blocks per SM: the row is held as the PACKED
// bf16 words that were loaded (d, y: 8 registers per 256-column slab instead of 16 floats) and unpacked again in the
// second phase; gamma is re-read from L1 per row.  Only the three column-sum sets stay in registers across rows, so
// the row tensors are read from HBM once (the rows + cols pair reads d and y twice and dz once more).
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
// bf16x2 word -> the two values as a packed fp32 pair (exact: a bf16 is the top half of an fp32)
__device__ __forceinline__ uint64_t unpack2(uint32_t w) {
  return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack2(uint64_t v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
// Instruction diet (the first version issued ~42 instructions per element and was issue-bound at 0.48 of the HBM peak):
// all arithmetic on packed fp32 pairs (FFMA2 / FMUL2 / FADD2), xhat as ONE fma (y * rstd - mean * rstd), the output as
// two (d * (gamma * rstd) - rstd * m1, then - rstd * m2 * xhat), gamma as fp32 in shared memory (LDS.128, no unpack),
// dropout applied to the PACKED bf16 result with an AND mask looked up per 8 keep-bits (256 x 16 B table in shared
// memory) instead of eight bit tests + selects; dbias sums the bf16 values that are stored (what the consumer reads).
template <int NV, bool HAS_DX2>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
ln_bwd_lean_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
                   const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const bf16* __restrict__ gamma, bf16* __restrict__ dy, int64_t lddy,
                   bf16* __restrict__ dz, int64_t lddz, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int H) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[ROW_WARPS][256];
  __shared__ __align__(16) float sG[NV * 256];
  __shared__ __align__(16) uint32_t sKeep[256][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < NV * 256; c += ROW_WARPS * 32) sG[c] = c < H ? __bfloat162float(gamma[c]) : 0.0f;
  {
    const uint32_t b = threadIdx.x;          // 256 threads <-> the 256 patterns of 8 keep-bits
#pragma unroll
    for (int k = 0; k < 4; ++k)
      sKeep[b][k] = (((b >> (2 * k)) & 1u) ? 0x0000ffffu : 0u) | (((b >> (2 * k + 1)) & 1u) ? 0xffff0000u : 0u);
  }
  __syncthreads();
  uint64_t ag[NV][4], ab[NV][4], az[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) { ag[i][k] = 0ull; ab[i][k] = 0ull; az[i][k] = 0ull; }
  const float invH = 1.0f / static_cast<float>(H);
  const uint64_t ds2 = pk2(dscale, dscale);

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + warp; row < M;
       row += static_cast<int64_t>(gridDim.x) * ROW_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    const float nmr = -mu * rs;
    const uint64_t rs2 = pk2(rs, rs), nmr2 = pk2(nmr, nmr);
    uint4 dw[NV], yw[NV];      // packed bf16 rows of dx and y, kept across the two phases
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      dw[i] = make_uint4(0u, 0u, 0u, 0u);
      yw[i] = make_uint4(0u, 0u, 0u, 0u);
      if (col < H) {
        dw[i] = *reinterpret_cast<const uint4*>(dx + row * lddx + col);
        yw[i] = *reinterpret_cast<const uint4*>(y + row * ldy + col);
      }
    }
    uint64_t s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        const uint32_t da[4] = {dw[i].x, dw[i].y, dw[i].z, dw[i].w};
        const uint32_t ya[4] = {yw[i].x, yw[i].y, yw[i].z, yw[i].w};
        uint32_t xa[4] = {0u, 0u, 0u, 0u};
        if (HAS_DX2) {
          // d = dx + dx2 is formed in fp32; re-packing the sum to bf16 would change the result, so the second
          // stream is not kept but re-read in phase 2 (an L1 hit)
          const uint4 t = *reinterpret_cast<const uint4*>(dx2 + row * lddx2 + col);
          xa[0] = t.x; xa[1] = t.y; xa[2] = t.z; xa[3] = t.w;
        }
        const float4 ga = *reinterpret_cast<const float4*>(sG + col), gb = *reinterpret_cast<const float4*>(sG + col + 4);
        const uint64_t g2[4] = {pk2(ga.x, ga.y), pk2(ga.z, ga.w), pk2(gb.x, gb.y), pk2(gb.z, gb.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t d2 = unpack2(da[k]);
          if (HAS_DX2) d2 = add2(d2, unpack2(xa[k]));
          const uint64_t xh2 = fma2(unpack2(ya[k]), rs2, nmr2);
          const uint64_t dg2 = mul2(d2, g2[k]);
          s1 = add2(s1, dg2);
          s2 = fma2(dg2, xh2, s2);
          ag[i][k] = fma2(d2, xh2, ag[i][k]);
          ab[i][k] = add2(ab[i][k], d2);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    upk2(s1, s1a, s1b);
    upk2(s2, s2a, s2b);
    const float m1 = warp_sum(s1a + s1b) * invH, m2 = warp_sum(s2a + s2b) * invH;
    const uint64_t nA2 = pk2(-rs * m1, -rs * m1), nB2 = pk2(-rs * m2, -rs * m2);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        const uint32_t da[4] = {dw[i].x, dw[i].y, dw[i].z, dw[i].w};
        const uint32_t ya[4] = {yw[i].x, yw[i].y, yw[i].z, yw[i].w};
        uint32_t xa[4] = {0u, 0u, 0u, 0u};
        if (HAS_DX2) {
          const uint4 t = *reinterpret_cast<const uint4*>(dx2 + row * lddx2 + col);
          xa[0] = t.x; xa[1] = t.y; xa[2] = t.z; xa[3] = t.w;
        }
        const float4 ga = *reinterpret_cast<const float4*>(sG + col), gb = *reinterpret_cast<const float4*>(sG + col + 4);
        const uint64_t g2[4] = {pk2(ga.x, ga.y), pk2(ga.z, ga.w), pk2(gb.x, gb.y), pk2(gb.z, gb.w)};
        uint32_t keep[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (dmask != nullptr) {
          const uint32_t w = __ldg(dmask + row * ldmask + (col >> 5));
          const uint4 kk = *reinterpret_cast<const uint4*>(sKeep[(w >> (col & 31)) & 0xFFu]);
          keep[0] = kk.x; keep[1] = kk.y; keep[2] = kk.z; keep[3] = kk.w;
        }
        uint32_t ow[4], zw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t d2 = unpack2(da[k]);
          if (HAS_DX2) d2 = add2(d2, unpack2(xa[k]));
          const uint64_t xh2 = fma2(unpack2(ya[k]), rs2, nmr2);
          const uint64_t o2 = fma2(nB2, xh2, fma2(d2, mul2(g2[k], rs2), nA2));
          ow[k] = pack2(o2);
          if (dmask != nullptr) {
            zw[k] = pack2(mul2(o2, ds2)) & keep[k];
            az[i][k] = add2(az[i][k], unpack2(zw[k]));
          } else {
            zw[k] = ow[k];
            az[i][k] = add2(az[i][k], o2);
          }
        }
        if (dy != nullptr) *reinterpret_cast<uint4*>(dy + row * lddy + col) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (dz != nullptr && (dmask != nullptr || dz != dy))
          *reinterpret_cast<uint4*>(dz + row * lddz + col) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
      }
    }
  }
  // block reduction of the three column-sum sets, 256 columns at a time, then one atomic per column
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 256 >= H) break;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo, hi;
        upk2(which == 0 ? ag[i][k] : (which == 1 ? ab[i][k] : az[i][k]), lo, hi);
        red[warp][lane * 8 + 2 * k] = lo;
        red[warp][lane * 8 + 2 * k + 1] = hi;
      }
      __syncthreads();
      const int c = threadIdx.x;  // 256 threads <-> 256 columns
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) t += red[w][c];
      if (i * 256 + c < H) atomicAdd(dst + i * 256 + c, t);
    }
  }
}

// Single-pass LayerNorm backward, streaming variant (MMFB_LN_BWD=stream).  ncu of the kernel above (round 2, call 6): issue
// slots 39 % used, 49 % of the stall samples on the global loads of the row a warp is about to process - a warp loads,
// computes, stores, and only the other 15 warps of the SM cover its load latency.  Here every warp owns a private ring of
// ST row slots in shared memory that its lane 0 fills with bulk async copies (cp.async.bulk, one per tensor row,
// completion on a per-slot mbarrier) ST rows ahead, so ST x 16 warps x 3 KB are in flight per SM whatever the warps are
// doing, the row lives in shared memory across the two phases (no row registers), and the arithmetic is the packed-fp32
// form of the lean kernel.  Rows are read from HBM once; dy / dz are stored straight from registers.
__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int NV, bool HAS_DX2, int ST>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
ln_bwd_stream_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
                     const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const bf16* __restrict__ gamma, bf16* __restrict__ dy, int64_t lddy,
                     bf16* __restrict__ dz, int64_t lddz, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int H) {
  griddep_launch();
  griddep_wait();
  constexpr int NT = HAS_DX2 ? 3 : 2;                 // tensors per slot: dx, y (, dx2)
  constexpr int ROWB = NV * 512;                      // bytes reserved per tensor row in a slot
  extern __shared__ __align__(128) uint8_t ln_smem[];
  uint8_t* ring = ln_smem;                                                        // [ROW_WARPS][ST][NT][ROWB]
  float* red = reinterpret_cast<float*>(ring + ROW_WARPS * ST * NT * ROWB);       // [ROW_WARPS][256]
  float* sG = red + ROW_WARPS * 256;                                              // [NV * 256]
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sG + NV * 256);                   // [256][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(sKeep + 1024);                     // [ROW_WARPS][ST]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < NV * 256; c += ROW_WARPS * 32) sG[c] = c < H ? __bfloat162float(gamma[c]) : 0.0f;
  {
    const uint32_t b = threadIdx.x;          // 256 threads <-> the 256 patterns of 8 keep-bits
#pragma unroll
    for (int k = 0; k < 4; ++k)
      sKeep[b * 4 + k] = (((b >> (2 * k)) & 1u) ? 0x0000ffffu : 0u) | (((b >> (2 * k + 1)) & 1u) ? 0xffff0000u : 0u);
  }
  if (threadIdx.x < ROW_WARPS * ST) mbar_init(&full[threadIdx.x], 1);
  fence_barrier_init();
  __syncthreads();
  uint8_t* my = ring + warp * (ST * NT * ROWB);
  uint64_t* my_full = full + warp * ST;
  const uint32_t row_bytes = static_cast<uint32_t>(H) * 2u;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * ROW_WARPS;
  const int64_t row_first = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + warp;
  auto fill = [&](int slot, int64_t row) {            // lane 0
    uint8_t* d = my + slot * (NT * ROWB);
    mbar_expect_tx(&my_full[slot], NT * row_bytes);
    bulk_load_1d(d, dx + row * lddx, row_bytes, &my_full[slot]);
    bulk_load_1d(d + ROWB, y + row * ldy, row_bytes, &my_full[slot]);
    if (HAS_DX2) bulk_load_1d(d + 2 * ROWB, dx2 + row * lddx2, row_bytes, &my_full[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s0 = 0; s0 < ST; ++s0)
      if (row_first + s0 * stride < M) fill(s0, row_first + s0 * stride);
  }
  uint64_t ag[NV][4], ab[NV][4], az[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) { ag[i][k] = 0ull; ab[i][k] = 0ull; az[i][k] = 0ull; }
  const float invH = 1.0f / static_cast<float>(H);
  const uint64_t ds2 = pk2(dscale, dscale);
  float mu_n = 0.0f, rs_n = 0.0f;
  if (row_first < M) { mu_n = mean[row_first]; rs_n = rstd[row_first]; }
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t row = row_first; row < M; row += stride) {
    const float mu = mu_n, rs = rs_n;
    if (row + stride < M) { mu_n = mean[row + stride]; rs_n = rstd[row + stride]; }   // one row ahead
    uint32_t keepw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      keepw[i] = 0xFFu;
      if (dmask != nullptr && col < H) keepw[i] = (__ldg(dmask + row * ldmask + (col >> 5)) >> (col & 31)) & 0xFFu;
    }
    const float nmr = -mu * rs;
    const uint64_t rs2 = pk2(rs, rs), nmr2 = pk2(nmr, nmr);
    const uint8_t* sl = my + slot * (NT * ROWB);
    mbar_wait(&my_full[slot], phase);
    uint64_t s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        const uint4 dwv = *reinterpret_cast<const uint4*>(sl + col * 2);
        const uint4 ywv = *reinterpret_cast<const uint4*>(sl + ROWB + col * 2);
        const uint32_t da[4] = {dwv.x, dwv.y, dwv.z, dwv.w};
        const uint32_t ya[4] = {ywv.x, ywv.y, ywv.z, ywv.w};
        uint32_t xa[4] = {0u, 0u, 0u, 0u};
        if (HAS_DX2) {
          const uint4 t = *reinterpret_cast<const uint4*>(sl + 2 * ROWB + col * 2);
          xa[0] = t.x; xa[1] = t.y; xa[2] = t.z; xa[3] = t.w;
        }
        const float4 ga = *reinterpret_cast<const float4*>(sG + col), gb = *reinterpret_cast<const float4*>(sG + col + 4);
        const uint64_t g2[4] = {pk2(ga.x, ga.y), pk2(ga.z, ga.w), pk2(gb.x, gb.y), pk2(gb.z, gb.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t d2 = unpack2(da[k]);
          if (HAS_DX2) d2 = add2(d2, unpack2(xa[k]));
          const uint64_t xh2 = fma2(unpack2(ya[k]), rs2, nmr2);
          const uint64_t dg2 = mul2(d2, g2[k]);
          s1 = add2(s1, dg2);
          s2 = fma2(dg2, xh2, s2);
          ag[i][k] = fma2(d2, xh2, ag[i][k]);
          ab[i][k] = add2(ab[i][k], d2);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    upk2(s1, s1a, s1b);
    upk2(s2, s2a, s2b);
    const float m1 = warp_sum(s1a + s1b) * invH, m2 = warp_sum(s2a + s2b) * invH;
    const uint64_t nA2 = pk2(-rs * m1, -rs * m1), nB2 = pk2(-rs * m2, -rs * m2);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = i * 256 + lane * 8;
      if (col < H) {
        const uint4 dwv = *reinterpret_cast<const uint4*>(sl + col * 2);
        const uint4 ywv = *reinterpret_cast<const uint4*>(sl + ROWB + col * 2);
        const uint32_t da[4] = {dwv.x, dwv.y, dwv.z, dwv.w};
        const uint32_t ya[4] = {ywv.x, ywv.y, ywv.z, ywv.w};
        uint32_t xa[4] = {0u, 0u, 0u, 0u};
        if (HAS_DX2) {
          const uint4 t = *reinterpret_cast<const uint4*>(sl + 2 * ROWB + col * 2);
          xa[0] = t.x; xa[1] = t.y; xa[2] = t.z; xa[3] = t.w;
        }
        const float4 ga = *reinterpret_cast<const float4*>(sG + col), gb = *reinterpret_cast<const float4*>(sG + col + 4);
        const uint64_t g2[4] = {pk2(ga.x, ga.y), pk2(ga.z, ga.w), pk2(gb.x, gb.y), pk2(gb.z, gb.w)};
        uint32_t keep[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (dmask != nullptr) {
          const uint4 kk = *reinterpret_cast<const uint4*>(sKeep + keepw[i] * 4);
          keep[0] = kk.x; keep[1] = kk.y; keep[2] = kk.z; keep[3] = kk.w;
        }
        uint32_t ow[4], zw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t d2 = unpack2(da[k]);
          if (HAS_DX2) d2 = add2(d2, unpack2(xa[k]));
          const uint64_t xh2 = fma2(unpack2(ya[k]), rs2, nmr2);
          const uint64_t o2 = fma2(nB2, xh2, fma2(d2, mul2(g2[k], rs2), nA2));
          ow[k] = pack2(o2);
          if (dmask != nullptr) {
            zw[k] = pack2(mul2(o2, ds2)) & keep[k];
            az[i][k] = add2(az[i][k], unpack2(zw[k]));
          } else {
            zw[k] = ow[k];
            az[i][k] = add2(az[i][k], o2);
          }
        }
        if (dy != nullptr) *reinterpret_cast<uint4*>(dy + row * lddy + col) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (dz != nullptr && (dmask != nullptr || dz != dy))
          *reinterpret_cast<uint4*>(dz + row * lddz + col) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
      }
    }
    // the slot has been read by every lane: refill it with the row ST steps ahead
    __syncwarp();
    if (lane == 0 && row + ST * stride < M) {
      fence_proxy_async();
      fill(slot, row + ST * stride);
    }
    if (++slot == ST) { slot = 0; phase ^= 1u; }
  }
  // block reduction of the three column-sum sets, 256 columns at a time, then one atomic per column
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 256 >= H) break;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo, hi;
        upk2(which == 0 ? ag[i][k] : (which == 1 ? ab[i][k] : az[i][k]), lo, hi);
        red[warp * 256 + lane * 8 + 2 * k] = lo;
        red[warp * 256 + lane * 8 + 2 * k + 1] = hi;
      }
      __syncthreads();
      const int c = threadIdx.x;  // 256 threads <-> 256 columns
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) t += red[w * 256 + c];
      if (i * 256 + c < H) atomicAdd(dst + i * 256 + c, t);
    }
  }
}

// Single-pass LayerNorm backward, tile variant (staged: MMFB_LN_BWD=tile).  A row is spread over WPR warps (each lane
// owns ONE 8-column slab for all rows it visits), so the three column-sum sets cost 24 registers per thread instead of
// 24 x (H / 256), gamma is loop-invariant, and a group of WPR warps processes RB rows per step: 2 x RB 16-byte loads in
// flight per lane and one named barrier per RB rows (row statistics of the WPR warps exchanged through shared memory,
// double-buffered by step parity).  dx, y, keep-bits are read once, dy / dz written once: traffic = the algorithmic bytes.
template <int WPR, int G, bool HAS_DX2>
__global__ void __launch_bounds__(WPR * G * 32, 1)
ln_bwd_tile_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
                   const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const bf16* __restrict__ gamma, bf16* __restrict__ dy, int64_t lddy,
                   bf16* __restrict__ dz, int64_t lddz, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int H) {
  griddep_launch();
  griddep_wait();
  constexpr int RB = 4;
  __shared__ float sStat[2][G][WPR][RB][2];
  __shared__ float sRed[G][WPR * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / WPR, wig = warp % WPR;
  const int col = (wig * 32 + lane) * 8;
  const bool cok = col < H;
  float g[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) g[e] = 0.0f;
  if (cok) ld8(gamma + col, g);
  float ag[8], ab[8], az[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { ag[e] = 0.0f; ab[e] = 0.0f; az[e] = 0.0f; }
  const float invH = 1.0f / static_cast<float>(H);
  const int64_t n_steps = (static_cast<int64_t>(M) + RB - 1) / RB;
  int par = 0;
  const int64_t g_first = static_cast<int64_t>(blockIdx.x) * G + grp, g_stride = static_cast<int64_t>(gridDim.x) * G;
  for (int64_t st = g_first; st < n_steps; st += g_stride, par ^= 1) {
    const int64_t row0 = st * RB;
    uint4 dw[RB], yw[RB];          // the rows as the packed bf16 words that were loaded; unpacked in both phases
    float mu[RB], rs[RB], s1[RB], s2[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int64_t row = row0 + r;
      dw[r] = make_uint4(0, 0, 0, 0);
      yw[r] = make_uint4(0, 0, 0, 0);
      if (cok && row < M) {
        dw[r] = *reinterpret_cast<const uint4*>(dx + row * lddx + col);
        yw[r] = *reinterpret_cast<const uint4*>(y + row * ldy + col);
      }
      mu[r] = row < M ? mean[row] : 0.0f;
      rs[r] = row < M ? rstd[row] : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float d[8], yy[8];
      unpack8(dw[r], d);
      if (HAS_DX2 && cok && row0 + r < M) {
        // d = dx + dx2 in fp32; the second stream is re-read in phase 2 (an L1 hit) rather than kept
        float t[8];
        ld8(dx2 + (row0 + r) * lddx2 + col, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
      }
      unpack8(yw[r], yy);
      float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (yy[e] - mu[r]) * rs[r];
        const float dg = d[e] * g[e];
        a1 += dg;
        a2 += dg * xh;
        ag[e] += d[e] * xh;
        ab[e] += d[e];
      }
      s1[r] = warp_sum(a1);
      s2[r] = warp_sum(a2);
    }
    if (WPR > 1) {
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RB; ++r) { sStat[par][grp][wig][r][0] = s1[r]; sStat[par][grp][wig][r][1] = s2[r]; }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) { t1 += sStat[par][grp][w][r][0]; t2 += sStat[par][grp][w][r][1]; }
        s1[r] = t1; s2[r] = t2;
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int64_t row = row0 + r;
      if (!(cok && row < M)) continue;
      float d[8], yy[8], o[8];
      unpack8(dw[r], d);
      if (HAS_DX2) {
        float t[8];
        ld8(dx2 + row * lddx2 + col, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
      }
      unpack8(yw[r], yy);
      const float m1 = s1[r] * invH, m2 = s2[r] * invH;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs[r] * (d[e] * g[e] - m1 - ((yy[e] - mu[r]) * rs[r]) * m2);
      if (dy != nullptr) st8(dy + row * lddy + col, o);
      if (dmask != nullptr) {
        const uint32_t w = __ldg(dmask + row * ldmask + (col >> 5));
        const uint32_t bits = (w >> (col & 31)) & 0xFFu;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ((bits >> e) & 1u) ? o[e] * dscale : 0.0f;
        st8(dz + row * lddz + col, o);
      } else if (dz != nullptr && dz != dy) {
        st8(dz + row * lddz + col, o);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) az[e] += o[e];
    }
  }
  // block reduction over the G groups, then one atomic per column and statistic
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) sRed[grp][(wig * 32 + lane) * 8 + e] = which == 0 ? ag[e] : (which == 1 ? ab[e] : az[e]);
    __syncthreads();
    for (int c = threadIdx.x; c < WPR * 256; c += WPR * G * 32) {
      float t = 0.0f;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) t += sRed[gg][c];
      if (c < H) atomicAdd(dst + c, t);
    }
  }
}

// Row-only LayerNorm backward (no column statistics): few registers, high occupancy.  dy, dz as above.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_bwd_rows_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
                   const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const bf16* __restrict__ gamma, bf16* __restrict__ dy, int64_t lddy,
                   bf16* __restrict__ dz, int64_t lddz, const uint32_t* __restrict__ dmask, int64_t ldmask, float dscale,
                   int M, int H) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  float dg[NV][8], xh[NV][8];
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      float d[8], yy[8], g[8];
      ld8(dx + row * lddx + col, d);
      if (dx2 != nullptr) {
        float t[8];
        ld8(dx2 + row * lddx2 + col, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
      }
      ld8(y + row * ldy + col, yy);
      ld8(gamma + col, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[i][e] = (yy[e] - mu) * rs;
        dg[i][e] = d[e] * g[e];
        s1 += dg[i][e];
        s2 += dg[i][e] * xh[i][e];
      }
    }
  }
  s1 = warp_sum(s1) / H;
  s2 = warp_sum(s2) / H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs * (dg[i][e] - s1 - xh[i][e] * s2);
      if (dy != nullptr) st8(dy + row * lddy + col, o);
      if (dmask != nullptr) {
        const uint32_t w = __ldg(dmask + row * ldmask + (col >> 5));
        const uint32_t bits = (w >> (col & 31)) & 0xFFu;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ((bits >> e) & 1u) ? o[e] * dscale : 0.0f;
        st8(dz + row * lddz + col, o);
      } else if (dz != nullptr && dz != dy) {
        st8(dz + row * lddz + col, o);
      }
    }
  }
}

// Column statistics of the LayerNorm backward: dgamma += sum_r d*xhat, dbeta += sum_r d, dbias += sum_r dz, with
// d = dx (+dx2).  Same access pattern as colsum: each thread owns 8 columns for a strided set of rows.
__global__ void __launch_bounds__(256)
ln_bwd_cols_kernel(const bf16* __restrict__ dx, int64_t lddx, const bf16* __restrict__ dx2, int64_t lddx2,
                   const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const bf16* __restrict__ dz, int64_t lddz,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int H) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float ag[8], ab[8], az[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { ag[e] = 0.0f; ab[e] = 0.0f; az[e] = 0.0f; }
  if (col < H) {
    for (int64_t row = static_cast<int64_t>(blockIdx.y) * 8 + warp; row < M; row += static_cast<int64_t>(gridDim.y) * 8) {
      const float mu = mean[row], rs = rstd[row];
      float d[8], yy[8];
      ld8(dx + row * lddx + col, d);
      if (dx2 != nullptr) {
        float t[8];
        ld8(dx2 + row * lddx2 + col, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += t[e];
      }
      ld8(y + row * ldy + col, yy);
#pragma unroll
      for (int e = 0; e < 8; ++e) { ag[e] += d[e] * (yy[e] - mu) * rs; ab[e] += d[e]; }
      if (dbias != nullptr) {
        float z[8];
        ld8(dz + row * lddz + col, z);
#pragma unroll
        for (int e = 0; e < 8; ++e) az[e] += z[e];
      }
    }
  }
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = which == 0 ? ag[e] : (which == 1 ? ab[e] : az[e]);
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < H) atomicAdd(dst + c, t);
  }
}

// ----------------------------------------------------------------------------------------------
// column sums (bias gradients):  out[n] += sum_m X[m, n]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ X, int64_t ldx, float* __restrict__ out, int M, int N) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  if (col < N) {
    // four independent 16-byte loads in flight per lane (the one-load-per-iteration loop reached 0.57 of the HBM peak),
    // packed fp32 adds
    uint64_t a2[4] = {0ull, 0ull, 0ull, 0ull};
    const int64_t stride = static_cast<int64_t>(gridDim.y) * 8;
    for (int64_t row = static_cast<int64_t>(blockIdx.y) * 8 + warp; row < M; row += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t r = row + k * stride;
        u[k] = r < M ? *reinterpret_cast<const uint4*>(X + r * ldx + col) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a2[0] = add2(a2[0], unpack2(u[k].x));
        a2[1] = add2(a2[1], unpack2(u[k].y));
        a2[2] = add2(a2[2], unpack2(u[k].z));
        a2[3] = add2(a2[3], unpack2(u[k].w));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) upk2(a2[k], acc[2 * k], acc[2 * k + 1]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) atomicAdd(out + c, t);
}

// ----------------------------------------------------------------------------------------------
// dropout keep-bits: word w, bit j keeps element 32*w + j with probability 1-p (16-bit resolution)
// ----------------------------------------------------------------------------------------------
// Bit-sliced comparison: the 16 Philox words of a thread are read as 16 bit-PLANES of 32 sixteen-bit uniforms r_e (bit e of
// plane b = bit b of r_e), and r_e >= thresh16 is evaluated for the 32 elements at once, most significant plane first
// (<= 2 logic operations per plane; the branch on the threshold bit is uniform).  Same distribution and resolution as
// comparing 16-bit fields one by one (the round-1 form: ~3 extract / compare / insert operations per ELEMENT).
// `epoch` (optional, device memory): a step counter added to the Philox counter as epoch << 40, so that a CUDA-graph replay of a
// captured step (host arguments frozen at capture) still draws fresh masks once the graph itself increments the counter.
__global__ void dropout_bits_kernel(uint32_t* __restrict__ out, int64_t nwords, uint64_t seed, uint64_t offset,
                                    uint32_t thresh16, const unsigned long long* __restrict__ epoch) {
  griddep_launch();
  griddep_wait();
  const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  if (epoch != nullptr) offset += static_cast<uint64_t>(*epoch) << 40;
  uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 r = philox4x32(seed, offset + static_cast<uint64_t>(w) * 4 + i);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = 15 - (i * 4 + k);              // plane of bit b, most significant first
      if ((thresh16 >> b) & 1u) {
        eq &= rr[k];                               // threshold bit 1: equal so far only where the uniform's bit is 1 too
      } else {
        gt |= eq & rr[k];                          // threshold bit 0, uniform bit 1: greater from here on
        eq &= ~rr[k];
      }
    }
  }
  out[w] = gt | eq;                                // keep where r >= thresh16
}

// ----------------------------------------------------------------------------------------------
// embedding compose:  y[r] = src0[s0[r]] + src1[s1[r]] + tab0[i0[r]] + tab1[i1[r]] + tab2[i2[r]]   (absent terms: index < 0)
// followed by LayerNorm (+dropout) -> x[r]; y (pre-LN) is saved for the backward.
// ----------------------------------------------------------------------------------------------
struct ComposeDev {
  const bf16* src[2]; int64_t ldsrc[2]; const int32_t* srow[2];
  const bf16* tab[3]; const int32_t* tidx[3];
};

__global__ void __launch_bounds__(ROW_WARPS * 32)
compose_kernel(ComposeDev c, bf16* __restrict__ y, int64_t ldy, int M, int H) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  int sidx[2], tix[3];
#pragma unroll
  for (int k = 0; k < 2; ++k) sidx[k] = (c.src[k] != nullptr && c.srow[k] != nullptr) ? c.srow[k][row] : -1;
#pragma unroll
  for (int k = 0; k < 3; ++k) tix[k] = (c.tab[k] != nullptr && c.tidx[k] != nullptr) ? c.tidx[k][row] : -1;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = 0.0f;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (sidx[k] >= 0) {
          float t[8];
          ld8(c.src[k] + static_cast<int64_t>(sidx[k]) * c.ldsrc[k] + col, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] += t[e];
        }
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (tix[k] >= 0) {
          float t[8];
          ld8(c.tab[k] + static_cast<int64_t>(tix[k]) * H + col, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] += t[e];
        }
      st8(y + row * ldy + col, a);
    }
  }
}

struct ScatterDev {
  bf16* dsrc[2]; int64_t ldsrc[2]; const int32_t* srow[2];
  float* dtab[3]; const int32_t* tidx[3];
};

// backward of compose: dsrc_k[s_k[r]] = dy[r]  (rows are unique) ; dtab_k[i_k[r]] += dy[r]  (fp32 atomics)
__global__ void __launch_bounds__(ROW_WARPS * 32)
scatter_kernel(ScatterDev c, const bf16* __restrict__ dy, int64_t lddy, int M, int H) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  int sidx[2], tix[3];
#pragma unroll
  for (int k = 0; k < 2; ++k) sidx[k] = (c.dsrc[k] != nullptr && c.srow[k] != nullptr) ? c.srow[k][row] : -1;
#pragma unroll
  for (int k = 0; k < 3; ++k) tix[k] = (c.dtab[k] != nullptr && c.tidx[k] != nullptr) ? c.tidx[k][row] : -1;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int col = i * 256 + lane * 8;
    if (col < H) {
      float a[8];
      ld8(dy + row * lddy + col, a);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (sidx[k] >= 0) st8(c.dsrc[k] + static_cast<int64_t>(sidx[k]) * c.ldsrc[k] + col, a);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (tix[k] >= 0) {
          float* d = c.dtab[k] + static_cast<int64_t>(tix[k]) * H + col;
#pragma unroll
          for (int e = 0; e < 8; ++e) atomicAdd(d + e, a[e]);
        }
    }
  }
}

// Table-gradient scatter over SORTED indices: warp w owns sorted positions [32w, 32w+32); rows with equal index are
// summed in registers and flushed with one atomic per column per run, so an index shared by thousands of rows (type /
// position embeddings) costs M/32 atomics per address instead of M.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
scatter_sorted_kernel(const bf16* __restrict__ dy, int64_t lddy, const int32_t* __restrict__ order,
                      const int32_t* __restrict__ sorted_idx, float* __restrict__ dtab, int M, int H) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * ROW_WARPS + (threadIdx.x >> 5);
  const int64_t p0 = w * 32;
  if (p0 >= M) return;
  float acc[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.0f;
  int cur = -1;
  const int n = (M - p0) < 32 ? static_cast<int>(M - p0) : 32;
  for (int k = 0; k <= n; ++k) {
    const int id = (k < n) ? sorted_idx[p0 + k] : -2;
    if (id != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int col = i * 256 + lane * 8;
          if (col < H) {
            float* d = dtab + static_cast<int64_t>(cur) * H + col;
#pragma unroll
            for (int e = 0; e < 8; ++e) { atomicAdd(d + e, acc[i][e]); acc[i][e] = 0.0f; }
          }
        }
      }
      cur = id;
    }
    if (k < n && id >= 0) {
      const int64_t r = order[p0 + k];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = i * 256 + lane * 8;
        if (col < H) {
          float t[8];
          ld8(dy + r * lddy + col, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] += t[e];
        }
      }
    }
  }
}

// backward of the ReLU epilogue: dz = (y > 0) ? dy : 0, 16-byte vectors
__global__ void relu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, bf16* __restrict__ dz, int64_t n) {
  griddep_launch();
  griddep_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float a[8], b[8];
    ld8(dy + i, a);
    ld8(y + i, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = b[e] > 0.0f ? a[e] : 0.0f;
    st8(dz + i, a);
  } else {
    for (int64_t j = i; j < n; ++j) dz[j] = __bfloat162float(y[j]) > 0.0f ? dy[j] : __float2bfloat16_rn(0.0f);
  }
}

// out = a + b (bf16, 16-byte vectors): the residual-gradient join of the pre-LN (ViT) layer, mmf/modules/vit.py:96-108
__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int64_t n) {
  griddep_launch();
  griddep_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float x[8], y[8];
    ld8(a + i, x);
    ld8(b + i, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    st8(out + i, x);
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(__bfloat162float(a[j]) + __bfloat162float(b[j]));
  }
}

// out[m, h] = keep-bit(m, h) ? x[m, h] * scale : 0 : the backward of a dropout that is NOT followed by a LayerNorm
// (whose backward kernel otherwise applies the mask): one warp per row, 8 columns per lane and step
__global__ void dropout_apply_kernel(const bf16* __restrict__ x, int64_t ldx, const uint32_t* __restrict__ bits, int64_t ldm,
                                     float scale, bf16* __restrict__ out, int64_t ldo, int M, int H) {
  griddep_launch();
  griddep_wait();
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  for (int col = lane * 8; col < H; col += 256) {
    float v[8];
    ld8(x + row * ldx + col, v);
    const uint32_t w = __ldg(bits + row * ldm + (col >> 5));
    const uint32_t b8 = (w >> (col & 31)) & 0xFFu;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = ((b8 >> e) & 1u) ? v[e] * scale : 0.0f;
    st8(out + row * ldo + col, v);
  }
}

// fp32 -> bf16 cast of the flat parameter buffer (done every forward, like autocast does)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  griddep_launch();
  griddep_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i);
    const float4 b = *reinterpret_cast<const float4*>(in + i + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + i) = u;
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

// ----------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------
static int check_rows(int M, int H, const char* who) {
  if (M <= 0 || H <= 0) return set_error(MMFB_ERR_ARG, "%s: empty problem", who);
  if (H % 8 || H > MAXV * 256) return set_error(MMFB_ERR_ARG, "%s: hidden size %d must be a multiple of 8 and <= %d", who, H, MAXV * 256);
  return MMFB_OK;
}
static int launch_ok(const char* who) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MMFB_ERR_CUDA, "%s launch: %s", who, cudaGetErrorString(e));
  count_launch();
  return MMFB_OK;
}

int ln_fwd(const mmfb_ln_args& a, cudaStream_t s) {
  int rc = check_rows(a.M, a.H, "layernorm_fwd");
  if (rc) return rc;
  if (!a.y || !a.gamma || !a.beta || !a.x) return set_error(MMFB_ERR_ARG, "layernorm_fwd: null pointer");
#define LN_FWD(NV)                                                                                         \
  MMFB_LAUNCH(ln_fwd_kernel<NV>, (a.M + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s,                           \
      (const bf16*)a.y, a.ldy, (const bf16*)a.gamma, (const bf16*)a.beta, (bf16*)a.x, a.ldx, a.mean, a.rstd, \
      a.drop_mask, a.ldmask, a.drop_scale, a.M, a.H, a.eps)
  const int nv = (a.H + 255) / 256;
  if (nv <= 1) LN_FWD(1); else if (nv == 2) LN_FWD(2); else if (nv == 3) LN_FWD(3); else if (nv == 4) LN_FWD(4); else LN_FWD(8);
#undef LN_FWD
  return launch_ok("layernorm_fwd");
}

int ln_bwd(const mmfb_ln_args& a, cudaStream_t s) {
  int rc = check_rows(a.M, a.H, "layernorm_bwd");
  if (rc) return rc;
  if (!a.dx || !a.y || !a.mean || !a.rstd || !a.gamma) return set_error(MMFB_ERR_ARG, "layernorm_bwd: null pointer");
  if (a.drop_mask && !a.dz) return set_error(MMFB_ERR_ARG, "layernorm_bwd: dropout mask given without dz");
  const int nv_ = (a.H + 255) / 256;
  // read per call (not cached) so that one test process can run the variants back to back.
  // default: the streaming single-pass kernel (58 us at [37848, 768] against 66 us for "lean", 80 us for "tile" and 100 us
  // for the rows + cols "pair", profiles/r2_kbench_persistent_attn_stream_ln.json); MMFB_LN_BWD=lean / tile / pair select
  // the others for A/B runs
  const char* lean_env = getenv("MMFB_LN_BWD");
  const bool lean = lean_env != nullptr && lean_env[0] == 'l';
  const bool tile = lean_env != nullptr && lean_env[0] == 't';
  if (tile && nv_ <= 4) {
    const int64_t n_steps = (static_cast<int64_t>(a.M) + 3) / 4;
#define LN_TILE(WPR, G)                                                                                                \
  {                                                                                                                    \
    int64_t want = (n_steps + (G) - 1) / (G);                                                                          \
    int grid = static_cast<int>(want < num_sms() ? want : num_sms());                                                  \
    if (grid < 1) grid = 1;                                                                                            \
    if (a.dx2 != nullptr)                                                                                              \
      MMFB_LAUNCH((ln_bwd_tile_kernel<WPR, G, true>), grid, (WPR) * (G) * 32, 0, s,                                               \
          (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2, (const bf16*)a.y, a.ldy, a.mean, a.rstd,             \
          (const bf16*)a.gamma, (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask, a.ldmask, a.drop_scale,         \
          a.dgamma, a.dbeta, a.dbias, a.M, a.H);                                                                       \
    else                                                                                                               \
      MMFB_LAUNCH((ln_bwd_tile_kernel<WPR, G, false>), grid, (WPR) * (G) * 32, 0, s,                                              \
          (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2, (const bf16*)a.y, a.ldy, a.mean, a.rstd,             \
          (const bf16*)a.gamma, (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask, a.ldmask, a.drop_scale,         \
          a.dgamma, a.dbeta, a.dbias, a.M, a.H);                                                                       \
  }
    if (nv_ <= 1) LN_TILE(1, 12) else if (nv_ == 2) LN_TILE(2, 6) else if (nv_ == 3) LN_TILE(3, 4) else LN_TILE(4, 3)
#undef LN_TILE
    return launch_ok("layernorm_bwd(tile)");
  }
  const bool stream = lean_env == nullptr || lean_env[0] == 's';
  // bulk copies need 16-byte aligned rows (H % 8 == 0 is already required; leading dimensions in elements % 8 too)
  if (stream && nv_ <= 4 && a.lddx % 8 == 0 && a.ldy % 8 == 0 && (a.dx2 == nullptr || a.lddx2 % 8 == 0)) {
    int grid = (a.M + ROW_WARPS - 1) / ROW_WARPS;
    const int cap = num_sms() * 2;
    if (grid > cap) grid = cap;
#define LN_STREAM(NV, DX2_, ST_)                                                                                      \
  {                                                                                                                    \
    constexpr size_t smem_ = static_cast<size_t>(ROW_WARPS) * (ST_) * ((DX2_) ? 3 : 2) * (NV) * 512 +                  \
                             ROW_WARPS * 256 * 4 + (NV) * 256 * 4 + 4096 + ROW_WARPS * (ST_) * 8;                      \
    static bool attr_done = false;                                                                                     \
    if (!attr_done) {                                                                                                  \
      cudaFuncSetAttribute(ln_bwd_stream_kernel<NV, DX2_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_); \
      attr_done = true;                                                                                                \
    }                                                                                                                  \
    MMFB_LAUNCH((ln_bwd_stream_kernel<NV, DX2_, ST_>), grid, ROW_WARPS * 32, smem_, s, (const bf16*)a.dx, a.lddx,      \
                (const bf16*)a.dx2, a.lddx2, (const bf16*)a.y, a.ldy, a.mean, a.rstd, (const bf16*)a.gamma,            \
                (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask, a.ldmask, a.drop_scale, a.dgamma, a.dbeta,      \
                a.dbias, a.M, a.H);                                                                                    \
  }
    if (a.dx2 != nullptr) {
      if (nv_ <= 1) LN_STREAM(1, true, 3) else if (nv_ == 2) LN_STREAM(2, true, 3) else if (nv_ == 3) LN_STREAM(3, true, 2) else LN_STREAM(4, true, 2)
    } else {
      if (nv_ <= 1) LN_STREAM(1, false, 4) else if (nv_ == 2) LN_STREAM(2, false, 4) else if (nv_ == 3) LN_STREAM(3, false, 3) else LN_STREAM(4, false, 2)
    }
#undef LN_STREAM
    return launch_ok("layernorm_bwd(stream)");
  }
  if ((lean || stream) && nv_ <= 4) {
    int grid = (a.M + ROW_WARPS - 1) / ROW_WARPS;
    const int cap = num_sms() * 2;      // two resident blocks per SM, rows strided over them
    if (grid > cap) grid = cap;
#define LN_LEAN(NV)                                                                                                  \
  MMFB_LAUNCH((ln_bwd_lean_kernel<NV, DX2>), grid, ROW_WARPS * 32, 0, s, (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2,       \
                                                         (const bf16*)a.y, a.ldy, a.mean, a.rstd, (const bf16*)a.gamma, \
                                                         (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask,         \
                                                         a.ldmask, a.drop_scale, a.dgamma, a.dbeta, a.dbias, a.M, a.H)
#define LN_LEAN2(DX2_)                                                                                              \
  { constexpr bool DX2 = DX2_;                                                                                       \
    if (nv_ <= 1) LN_LEAN(1); else if (nv_ == 2) LN_LEAN(2); else if (nv_ == 3) LN_LEAN(3); else LN_LEAN(4); }
    if (a.dx2 != nullptr) LN_LEAN2(true) else LN_LEAN2(false)
#undef LN_LEAN2
#undef LN_LEAN
    return launch_ok("layernorm_bwd(lean)");
  }
  if (nv_ <= 4) {
    // two kernels: a light row kernel (dy, dz) at high occupancy, then coalesced column statistics.  (Measured
    // alternatives: one fused kernel with register accumulators - 188 registers, one block per SM, 51 us; one fused
    // kernel with red.shared column partials - 72 shared reductions per lane per row, slower end to end.  The pair: 23 + 22 us.)
    const bool need_dz_buf = (a.drop_mask != nullptr) || (a.dz != nullptr && a.dz != a.dy);
    const int rgrid = (a.M + ROW_WARPS - 1) / ROW_WARPS;
#define LN_ROWS(NV)                                                                                              \
  MMFB_LAUNCH(ln_bwd_rows_kernel<NV>, rgrid, ROW_WARPS * 32, 0, s, (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2, \
                                                          (const bf16*)a.y, a.ldy, a.mean, a.rstd, (const bf16*)a.gamma, \
                                                          (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask, a.ldmask, \
                                                          a.drop_scale, a.M, a.H)
    if (nv_ <= 1) LN_ROWS(1); else if (nv_ == 2) LN_ROWS(2); else if (nv_ == 3) LN_ROWS(3); else LN_ROWS(4);
#undef LN_ROWS
    rc = launch_ok("layernorm_bwd(rows)");
    if (rc) return rc;
    if (a.dgamma || a.dbeta || a.dbias) {
      const void* dzp = need_dz_buf ? a.dz : a.dy;
      const int64_t lddz = need_dz_buf ? a.lddz : a.lddy;
      if (a.dbias && !dzp) return set_error(MMFB_ERR_ARG, "layernorm_bwd: dbias needs dy or dz to be materialised");
      dim3 cgrid((a.H + 255) / 256, 1);
      int rb = (a.M + 63) / 64;
      const int cap2 = (num_sms() * 4 + cgrid.x - 1) / cgrid.x;
      cgrid.y = rb < cap2 ? rb : cap2;
      if (cgrid.y < 1) cgrid.y = 1;
      MMFB_LAUNCH(ln_bwd_cols_kernel, cgrid, 256, 0, s, (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2, (const bf16*)a.y, a.ldy,
                                               a.mean, a.rstd, (const bf16*)dzp, lddz, a.dgamma, a.dbeta, a.dbias, a.M, a.H);
      return launch_ok("layernorm_bwd(cols)");
    }
    return MMFB_OK;
  }
  int grid = (a.M + ROW_WARPS - 1) / ROW_WARPS;
  const int cap = num_sms() * 2;
  if (grid > cap) grid = cap;
#define LN_BWD(NV)                                                                                              \
  MMFB_LAUNCH(ln_bwd_kernel<NV>, grid, ROW_WARPS * 32, 0, s, (const bf16*)a.dx, a.lddx, (const bf16*)a.dx2, a.lddx2,       \
                                                    (const bf16*)a.y, a.ldy, a.mean, a.rstd, (const bf16*)a.gamma, \
                                                    (bf16*)a.dy, a.lddy, (bf16*)a.dz, a.lddz, a.drop_mask,         \
                                                    a.ldmask, a.drop_scale, a.dgamma, a.dbeta, a.dbias, a.M, a.H)
  const int nv = (a.H + 255) / 256;
  if (nv <= 1) LN_BWD(1); else if (nv == 2) LN_BWD(2); else if (nv == 3) LN_BWD(3); else if (nv == 4) LN_BWD(4); else LN_BWD(8);
#undef LN_BWD
  return launch_ok("layernorm_bwd");
}

int colsum(const void* X, int64_t ldx, float* out, int M, int N, cudaStream_t s) {
  if (M <= 0 || N <= 0 || N % 8) return set_error(MMFB_ERR_ARG, "colsum: bad shape %dx%d", M, N);
  dim3 grid((N + 255) / 256, 1);
  int rows_blocks = (M + 63) / 64;
  const int cap = (num_sms() * 4 + grid.x - 1) / grid.x;
  grid.y = rows_blocks < cap ? rows_blocks : cap;
  if (grid.y < 1) grid.y = 1;
  MMFB_LAUNCH(colsum_kernel, grid, 256, 0, s, (const bf16*)X, ldx, out, M, N);
  return launch_ok("colsum");
}

int dropout_bits(uint32_t* out, int64_t nwords, uint64_t seed, uint64_t offset, float p, const uint64_t* epoch, cudaStream_t s) {
  if (nwords <= 0) return set_error(MMFB_ERR_ARG, "dropout_bits: empty");
  if (!(p >= 0.0f && p < 1.0f)) return set_error(MMFB_ERR_ARG, "dropout_bits: p must be in [0,1), got %f", p);
  uint32_t thresh = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  if (thresh > 65535u) thresh = 65535u;          // 16 planes: p within 2^-16 of 1 keeps one element in 65536
  MMFB_LAUNCH(dropout_bits_kernel, static_cast<unsigned>((nwords + 255) / 256), 256, 0, s, out, nwords, seed, offset, thresh,
              reinterpret_cast<const unsigned long long*>(epoch));
  return launch_ok("dropout_bits");
}

int compose(const mmfb_compose_args& a, cudaStream_t s) {
  int rc = check_rows(a.M, a.H, "embed_compose");
  if (rc) return rc;
  if (!a.y) return set_error(MMFB_ERR_ARG, "embed_compose: null output");
  ComposeDev c;
  for (int k = 0; k < 2; ++k) { c.src[k] = (const bf16*)a.src[k]; c.ldsrc[k] = a.ldsrc[k]; c.srow[k] = a.src_row[k]; }
  for (int k = 0; k < 3; ++k) { c.tab[k] = (const bf16*)a.tab[k]; c.tidx[k] = a.tab_idx[k]; }
  MMFB_LAUNCH(compose_kernel, (a.M + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s, c, (bf16*)a.y, a.ldy, a.M, a.H);
  return launch_ok("embed_compose");
}

int scatter(const mmfb_scatter_args& a, cudaStream_t s) {
  int rc = check_rows(a.M, a.H, "embed_scatter");
  if (rc) return rc;
  if (!a.dy) return set_error(MMFB_ERR_ARG, "embed_scatter: null input");
  ScatterDev c;
  for (int k = 0; k < 2; ++k) { c.dsrc[k] = (bf16*)a.dsrc[k]; c.ldsrc[k] = a.ldsrc[k]; c.srow[k] = a.src_row[k]; }
  for (int k = 0; k < 3; ++k) { c.dtab[k] = a.dtab[k]; c.tidx[k] = a.tab_idx[k]; }
  MMFB_LAUNCH(scatter_kernel, (a.M + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s, c, (const bf16*)a.dy, a.lddy, a.M, a.H);
  return launch_ok("embed_scatter");
}

int scatter_sorted(const void* dy, int64_t lddy, const int32_t* order, const int32_t* sorted_idx, float* dtab, int M,
                   int H, cudaStream_t s) {
  int rc = check_rows(M, H, "embed_scatter_sorted");
  if (rc) return rc;
  if (H > 1024) return set_error(MMFB_ERR_ARG, "embed_scatter_sorted: hidden size %d > 1024", H);
  const int warps = (M + 31) / 32;
  const int grid = (warps + ROW_WARPS - 1) / ROW_WARPS;
  const int nv = (H + 255) / 256;
#define SS(NV) MMFB_LAUNCH(scatter_sorted_kernel<NV>, grid, ROW_WARPS * 32, 0, s, (const bf16*)dy, lddy, order, sorted_idx, dtab, M, H)
  if (nv <= 1) SS(1); else if (nv == 2) SS(2); else if (nv == 3) SS(3); else SS(4);
#undef SS
  return launch_ok("embed_scatter_sorted");
}

int relu_bwd(const void* dy, const void* y, void* dz, int64_t n, cudaStream_t s) {
  if (n <= 0) return set_error(MMFB_ERR_ARG, "relu_bwd: empty");
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dz)) & 15)
    return set_error(MMFB_ERR_ARG, "relu_bwd: buffers must be 16-byte aligned");
  const int64_t thr = (n + 7) / 8;
  MMFB_LAUNCH(relu_bwd_kernel, static_cast<unsigned>((thr + 255) / 256), 256, 0, s, (const bf16*)dy, (const bf16*)y, (bf16*)dz, n);
  return launch_ok("relu_bwd");
}

int add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return set_error(MMFB_ERR_ARG, "add_bf16: empty");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MMFB_ERR_ARG, "add_bf16: buffers must be 16-byte aligned");
  const int64_t thr = (n + 7) / 8;
  MMFB_LAUNCH(add_bf16_kernel, static_cast<unsigned>((thr + 255) / 256), 256, 0, s, (const bf16*)a, (const bf16*)b, (bf16*)out, n);
  return launch_ok("add_bf16");
}

int dropout_apply(const void* x, int64_t ldx, const uint32_t* bits, int64_t ldm, float scale, void* out, int64_t ldo, int M,
                  int H, cudaStream_t s) {
  int rc = check_rows(M, H, "dropout_apply");
  if (rc) return rc;
  const int64_t threads = static_cast<int64_t>(M) * 32;
  MMFB_LAUNCH(dropout_apply_kernel, static_cast<unsigned>((threads + 255) / 256), 256, 0, s, (const bf16*)x, ldx, bits, ldm, scale,
                                                                                   (bf16*)out, ldo, M, H);
  return launch_ok("dropout_apply");
}

int cast_params(const float* in, void* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return set_error(MMFB_ERR_ARG, "cast: empty");
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(MMFB_ERR_ARG, "cast: buffers must be 16-byte aligned");
  const int64_t thr = (n + 7) / 8;
  MMFB_LAUNCH(cast_f32_bf16_kernel, static_cast<unsigned>((thr + 255) / 256), 256, 0, s, in, (bf16*)out, n);
  return launch_ok("cast_f32_bf16");
}

}  // namespace mmfb
