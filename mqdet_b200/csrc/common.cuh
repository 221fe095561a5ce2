// mqdet_b200 — shared device/host helpers for the sm_100a kernels.
// PTX wrappers (mbarrier, TMA, tcgen05) are written out here so the kernels read as plain CUDA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MQDET_OK 0
#define MQDET_ERR_ARG -1
#define MQDET_ERR_CUDA -2
#define MQDET_ERR_UNSUPPORTED -3

namespace mqdet {

// thread-local last error string (exposed through mqdet_last_error()).
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define MQ_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      mqdet::set_error(__VA_ARGS__);               \
      return MQDET_ERR_ARG;                        \
    }                                              \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// All FPN levels of an image live in one tensor x[B][N][C] (level l at row offset off_l, row-major (h, w)).
#ifndef MQDET_MAX_LEVELS
#define MQDET_MAX_LEVELS 8
#endif
struct LevelTable {
  int n;
  int H[MQDET_MAX_LEVELS], W[MQDET_MAX_LEVELS], off[MQDET_MAX_LEVELS];
};
static inline int fill_levels(LevelTable* t, const int32_t* hw, int64_t nlev) {
  if (nlev < 1 || nlev > MQDET_MAX_LEVELS) return -1;
  t->n = (int)nlev;
  int off = 0;
  for (int l = 0; l < nlev; ++l) {
    t->H[l] = hw[2 * l];
    t->W[l] = hw[2 * l + 1];
    t->off[l] = off;
    off += t->H[l] * t->W[l];
  }
  return off;
}

// host helpers of the tcgen05 kernels (capi.cu): cached TMA tensor maps, per-device SM count / shared-memory opt-in
int make_operand_map(CUtensorMap* map, const void* ptr, long rows, long K, long ld, int nb1, long s1, int nb2, long s2,
                     int box_rows, int* bcast1, int* bcast2);
int make_store_map(CUtensorMap* map, void* C, int c_dtype, long M, long N, long ldc, int nb1, long c_b1, int nb2, long c_b2);
int num_sms();
int ensure_dyn_smem(const void* func, int bytes);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// erf-GELU in ~18 issue slots (no branches): erfc(|z|) = t * P4(t) * exp(-z^2), t = 1 / (1 + 0.3275911 |z|)
// (Abramowitz-Stegun 7.1.26, |abs err| < 1.5e-7), and 1 + erf(z) = erfc(|z|) for z < 0 (no cancellation), 2 - erfc(z) else.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.f)));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float e = q * t * __expf(-az * az);
  return 0.5f * x * (z < 0.f ? e : 2.f - e);
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// explicit shared-space 128-bit accesses on 32-bit addresses (pointers derived from the aligned dynamic-smem base are
// otherwise compiled as generic LD.E/ST.E with 64-bit address arithmetic)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU.EX2 (flushes results below 2^-126 to zero)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// packed fp32x2 FMA (sm_100): (d0, d1) = (a0, a1) * (s, s) + (b0, b1) in ONE issue slot
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float s, float b0, float b1) {
  asm("{\n"
      ".reg .b64 va, vs, vb, vd;\n"
      "mov.b64 va, {%2, %3};\n"
      "mov.b64 vs, {%4, %4};\n"
      "mov.b64 vb, {%5, %6};\n"
      "fma.rn.f32x2 vd, va, vs, vb;\n"
      "mov.b64 {%0, %1}, vd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(s), "f"(b0), "f"(b1));
}
// (d0, d1) = (a0, a1) * (s0, s1) + (b0, b1)
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float s0, float s1, float b0, float b1) {
  asm("{\n"
      ".reg .b64 va, vs, vb, vd;\n"
      "mov.b64 va, {%2, %3};\n"
      "mov.b64 vs, {%4, %5};\n"
      "mov.b64 vb, {%6, %7};\n"
      "fma.rn.f32x2 vd, va, vs, vb;\n"
      "mov.b64 {%0, %1}, vd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(s0), "f"(s1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// ---- TMA (cp.async.bulk.tensor) ------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load, coordinates innermost-first.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// L2 prefetch of a 4-D tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// 4-D tiled store shared -> global (bulk group completion); out-of-bounds parts of the box are clipped by the TMA unit.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read1() {  // leave the newest group in flight
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Whole warp. Writes the TMEM base address (lane<<16 | column) to *smem_out.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16/bf16 operands, fp32 accumulate).
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// wait::ld that also "touches" the destination registers of an earlier tcgen05.ld, so that the compiler cannot schedule
// their first use above the wait when the load was issued several statements earlier (software-pipelined epilogue)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (8-row groups 1024 B apart).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version = 1 (Blackwell)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major (the M / N index is the contiguous one), 128-byte-swizzled descriptor: the swizzle atom is 64 MN elements (128 B)
// x 8 K rows; `lbo` = byte distance between atoms along MN (next 64 elements), `sbo` = between atoms along K (next 8 rows).
// Used with the matching major bit of the instruction descriptor (UMMA_A_MN_MAJOR / UMMA_B_MN_MAJOR).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t UMMA_A_MN_MAJOR = 1u << 15;
constexpr uint32_t UMMA_B_MN_MAJOR = 1u << 16;

// warp-wide max of an fp32 value in one instruction (sm_100a CREDUX.MAX.F32); every lane receives the result
__device__ __forceinline__ float warp_redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// erf-GELU of TWO values with the packed fp32x2 FMA pipe (sm_100): the same Abramowitz-Stegun 7.1.26 form as gelu_fast,
//   gelu(x) = 0.5 (x + |x| (1 - erfc(|z|))),  z = x / sqrt(2),  erfc(|z|) = t P4(t) exp(-z^2),  t = 1 / (1 + 0.3275911 |z|),
// with z pre-scaled by sqrt(log2 e) so that exp(-z^2) is one MUFU.EX2 of -(z')^2.  ~10 issue slots per value instead of ~17:
// the GELU epilogues of the K <= 384 MLP GEMMs (Swin fc1, FFNs) are issue-bound on the eight epilogue warps.
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  constexpr float SL = 1.2011224087864498f;                 // sqrt(log2(e))
  constexpr float C = 0.70710678118654752440f * SL;          // x -> z' = x / sqrt(2) * sqrt(log2 e)
  constexpr float P = 0.3275911f / SL;
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  float z0, z1, d0, d1, q0, q1, w0, w1, e0, e1, u0, u1, h0, h1;
  ffma2(z0, z1, a0, a1, C, 0.f, 0.f);                         // |z'|
  ffma2(d0, d1, z0, z1, P, 1.f, 1.f);                         // 1 + 0.3275911 |z|
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  ffma2(q0, q1, t0, t1, 1.061405429f, -1.453152027f, -1.453152027f);
  ffma2v(q0, q1, q0, q1, t0, t1, 1.421413741f, 1.421413741f);
  ffma2v(q0, q1, q0, q1, t0, t1, -0.284496736f, -0.284496736f);
  ffma2v(q0, q1, q0, q1, t0, t1, 0.254829592f, 0.254829592f);
  ffma2v(w0, w1, z0, z1, z0, z1, 0.f, 0.f);                   // (z')^2 = z^2 log2 e
  e0 = ex2_approx(-w0);
  e1 = ex2_approx(-w1);
  ffma2v(q0, q1, q0, q1, t0, t1, 0.f, 0.f);                   // t P4(t)
  ffma2v(e0, e1, q0, q1, e0, e1, 0.f, 0.f);                   // erfc(|z|)
  ffma2v(u0, u1, -a0, -a1, e0, e1, a0, a1);                   // |x| (1 - erfc)
  ffma2(h0, h1, x0, x1, 0.5f, 0.f, 0.f);
  ffma2(x0, x1, u0, u1, 0.5f, h0, h1);                        // 0.5 x + 0.5 |x| (1 - erfc)
}

// kind::f16 instruction descriptor: fp16 A/B (format 0) or bf16 (1), fp32 accumulate, K-major A/B.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int ab_format) {
  return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace mqdet
