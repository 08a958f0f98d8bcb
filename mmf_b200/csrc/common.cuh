// Shared device-side primitives for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM
// wrappers (inline PTX), UMMA descriptor builders, Philox dropout, small math helpers.
//
// Everything in csrc/ is written for sm_100a only (B200); there is no other code path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mmfb {

typedef __nv_bfloat16 bf16;

// --------------------------------------------------------------------------------------
// misc
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// First 1024-byte aligned address of the dynamic shared-memory window (SW128 tiles need it).  Pure pointer arithmetic
// on the __shared__ array - an integer round trip would make every later access a generic LD.E/ST.E with 64-bit
// address arithmetic instead of LDS/STS.
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// --------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel of the library lets the NEXT kernel of the stream start launching as soon
// as all of its own CTAs have started (griddep_launch at the top), and waits for the PREVIOUS kernel to have completed and
// flushed its memory before it touches global data (griddep_wait after its own set-up: barrier init, TMEM allocation,
// descriptor prefetch).  The launch latency and the prologue of kernel n+1 overlap the tail of kernel n; both
// instructions are no-ops when the kernel was launched without the attribute (MMFB_PDL=0).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// --------------------------------------------------------------------------------------
// mbarrier
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (mbarrier.test_wait returns at once; try_wait may suspend the thread for a hardware time slice)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a protocol bug must abort the kernel (trap) instead of hanging the GPU.
// The spin loop is kept to three instructions per failed probe (try_wait itself suspends the warp in hardware for a
// while before it returns false): the clock is read only every 2^16 probes.  The round-1 loop read the clock on every
// probe, and ncu showed 27 % of the attention forward's issued warp-instructions in it (CS2R / ISETP / BRA), taken from
// the second CTA on the SM.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 0xFFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {  // ~4 s at 2 GHz
        printf("mmfb: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}

// --------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), loads signal an mbarrier with complete_tx::bytes
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// multicast variant: the box lands at the same CTA-relative smem offset in every CTA of `mask`, and each destination
// CTA's mbarrier (same CTA-relative address) receives the complete_tx bytes
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// --------------------------------------------------------------------------------------
// tcgen05 / TMEM
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// ---- cta_group::2 (CTA pair) variants --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows in each CTA's smem] * B[N/2 rows in each CTA's smem]; issued by the leader only
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// TMA load into THIS CTA's smem whose completion bytes are credited to the mbarrier at the same offset in the pair's
// leader CTA (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at this CTA-relative address in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// same, arriving on the barrier at this CTA-relative address in every CTA of `mask` (cluster pipelines)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (row) t, columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM, 32 lanes x 16 columns (thread t of the warp writes lane t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (K-major only: lane = row, 16-bit elements
// packed two per 32-bit column in K order, so one K=16 step spans 8 columns).  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// --------------------------------------------------------------------------------------
// UMMA shared-memory matrix descriptors (SWIZZLE_128B canonical layouts, bf16)
//
// K-major operand tile  [rows][64 k]  : each row is one 128-byte line (64 bf16 along K), lines of 8
//   consecutive rows form a 1024-byte swizzle atom (exactly what a TMA box {64, rows} with
//   CU_TENSOR_MAP_SWIZZLE_128B writes).  SBO = 1024 B (next 8 rows). One UMMA (K=16) reads 32 bytes
//   of each line, so the k-th UMMA of the tile starts 32*k bytes further.
// MN-major operand tile [64 k][64 mn] per 8 KB panel: each line is one k, 64 bf16 along M/N (TMA box
//   {64 mn, 64 k}); SBO = 1024 B (next 8 k), LBO = bytes between 64-wide mn panels.  One UMMA (K=16)
//   reads 16 lines = 2048 bytes.
// Descriptor bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=2.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor, kind::f16, A/B = bf16, D = fp32
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                              // D format: F32
         | (1u << 7)                            // A format: BF16
         | (1u << 10)                           // B format: BF16
         | ((a_mn_major ? 1u : 0u) << 15)       // A major
         | ((b_mn_major ? 1u : 0u) << 16)       // B major
         | (static_cast<uint32_t>(N >> 3) << 17)  // N
         | (static_cast<uint32_t>(M >> 4) << 24); // M
}

// --------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG (7 rounds: the smallest variant that passes BigCrush) for the dropout keep-bits
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32), c2 = 0x243F6A88u,
           c3 = 0x85A308D3u;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask for the 4 consecutive elements [4*quad, 4*quad+4): bit i set = keep element i.
// thresh = floor(p * 2^32); element kept iff rnd >= thresh.
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint64_t quad, uint32_t thresh) {
  uint4 r = philox4x32(seed, quad);
  return (r.x >= thresh ? 1u : 0u) | (r.y >= thresh ? 2u : 0u) | (r.z >= thresh ? 4u : 0u) |
         (r.w >= thresh ? 8u : 0u);
}

// --------------------------------------------------------------------------------------
// math
// --------------------------------------------------------------------------------------
// exact (erf) GELU, branch-free: erf(z) = sign(z) (1 - P(t) exp(-z^2)), t = 1/(1 + 0.3275911 |z|)  (Abramowitz-Stegun
// 7.1.26, |error| <= 1.5e-7, i.e. fp32 rounding level).  With z = x/sqrt(2), exp(-z^2) = exp(-x^2/2) is also the
// Gaussian density needed by the derivative, so forward and backward cost one ex2 + one rcp (both MUFU approx,
// no IEEE slow paths -> no branches) and ~15 FMA-pipe instructions per element.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_phi_parts(float x, float& e) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));          // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);                  // erf(|x|/sqrt(2))
  return fmaf(0.5f, copysignf(erf_abs, x), 0.5f);                  // Phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  return x * gelu_phi_parts(x, e);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float cdf = gelu_phi_parts(x, e);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE fp32 lanes per FMA-pipe issue slot) ----
// Lane results are bit-identical to the scalar fmaf / * / + (round-to-nearest, no ftz), so a packed epilogue is a pure
// issue-slot optimisation.  On by default since round 2 (isolated timings, profiles/r2_kbench_before*.json: FFN-up GEMM
// + GELU 165 -> 152 us, GELU' dgrad 177 -> 154 us, fused attention backward 307 -> 296 us); -DMMFB_F32X2=0 builds the
// scalar variant for A/B runs.  See gelu_erf2 / gelu_erf_grad_mul2.
#ifndef MMFB_F32X2
#define MMFB_F32X2 1
#endif
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// Phi(x) and exp(-x^2/2) for the pair (x0, x1): the arithmetic of gelu_phi_parts, two lanes at a time (the negated
// polynomial coefficients give -(poly) exactly, so -poly*t needs no separate negation)
__device__ __forceinline__ uint64_t gelu_phi_parts2(float x0, float x1, uint64_t x, uint64_t& e) {
  const uint64_t den = fma2(pk2(fabsf(x0), fabsf(x1)), pk2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f),
                            pk2(1.0f, 1.0f));
  float d0, d1;
  upk2(den, d0, d1);
  const uint64_t t = pk2(rcp_approx(d0), rcp_approx(d1));
  const uint64_t arg = mul2(mul2(x, x), pk2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  float a0, a1;
  upk2(arg, a0, a1);
  e = pk2(ex2_approx(a0), ex2_approx(a1));
  uint64_t np = fma2(pk2(-1.061405429f, -1.061405429f), t, pk2(1.453152027f, 1.453152027f));
  np = fma2(np, t, pk2(-1.421413741f, -1.421413741f));
  np = fma2(np, t, pk2(0.284496736f, 0.284496736f));
  np = fma2(np, t, pk2(-0.254829592f, -0.254829592f));
  const uint64_t erf_abs = fma2(mul2(np, t), e, pk2(1.0f, 1.0f));
  float r0, r1;
  upk2(erf_abs, r0, r1);
  return fma2(pk2(0.5f, 0.5f), pk2(copysignf(r0, x0), copysignf(r1, x1)), pk2(0.5f, 0.5f));
}
__device__ __forceinline__ void gelu_erf2(float x0, float x1, float& y0, float& y1) {
  const uint64_t x = pk2(x0, x1);
  uint64_t e;
  upk2(mul2(x, gelu_phi_parts2(x0, x1, x, e)), y0, y1);
}
// (v0, v1) *= GELU'(x0), GELU'(x1)
__device__ __forceinline__ void gelu_erf_grad_mul2(float x0, float x1, float& v0, float& v1) {
  const uint64_t x = pk2(x0, x1);
  uint64_t e;
  const uint64_t cdf = gelu_phi_parts2(x0, x1, x, e);
  const uint64_t g = fma2(mul2(x, pk2(0.3989422804014327f, 0.3989422804014327f)), e, cdf);
  upk2(mul2(pk2(v0, v1), g), v0, v1);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mmfb
