// mqdet_b200 — the bi-directional image <-> text attention of the VL fusion tower (BiMultiHeadAttention,
// maskrcnn_benchmark/utils/fuse_helper.py:218-303) as two tcgen05 kernels that never materialise the score matrix:
//
//   biattn_image_kernel : image -> text side.  Per 128 image tokens and head: S = Q K^T (fp32 in TMEM) -> +-5e4 clamp ->
//                         column-max partials (for the text side) -> masked row softmax over the T tokens in registers ->
//                         P (fp16, shared memory) -> O = P V_l -> fp16 -> out-projection accumulated over the heads in a second
//                         TMEM accumulator -> v' = residual + gamma * (acc + bias) -> TMA store.  Neither the scores A nor the
//                         per-head context reach HBM; the softmax statistics are taken from the fp32 accumulator.
//   biattn_text_kernel  : text -> image side (one CTA per (image, head, 128 text tokens), streams all image tokens):
//                         S^T recomputed on the tensor cores, exp(S^T - column max), P . V accumulated in TMEM.
//                         VN variant: V = the layer-normed image tokens themselves (read MN-major from the same [n x c] tile),
//                         the value projection is applied AFTER the token reduction (sum_n p[n] = 1), and the column sums are
//                         accumulated in-kernel, so the [B, E, N] value tensor of the image side is never built.
#include "common.cuh"
#include "../../include/mqdet_b200.h"

namespace mqdet {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row

// =====================================================================================================================
// Fused text -> image side of the bi-directional attention (BiMultiHeadAttention, fuse_helper.py:257-275, 289-291):
//     out[t, :] = sum_n softmax_n(clamp(k_t . q_n)) * Vv[n, :]        per (image, head), t = text token, n = image token
// One CTA per (z, 128 text tokens).  The transposed probabilities [T, N] never reach HBM.  Per step of 256 image tokens:
//     S^T   = K_tile (128 x 256, resident) . Q_step^T      16 tcgen05 MMAs 128x256x16, Q streamed as four [256 x 64] k-blocks
//     P_j   = exp(fp16(clamp(S^T[:, 64j..])) - colmax_t)   -> fp16, 128B-swizzled A-operand tile in shared memory (j = 0..3)
//     O    += P_j . Vv_j                                    4 MMAs 128x256x16 per 64-token sub-block, O (128 x 256) in TMEM
//     out   = O / colsum_t  -> fp16 -> TMA store
// Every MMA is 256 wide: narrower ones (the first version used 64-token steps) pay ~the same ~130-160 cycles per
// instruction for a fraction of the work.  colmax / colsum come from mqdet_colsoftmax_stats / mqdet_colstats_rowsoftmax over
// the SAME fp16 scores the image -> text side uses, so no online rescaling is needed.
// Roles: warp 0 = Q producer, warp 3 = V producer (separate rings: a Q k-block is free as soon as its MMAs retire),
// warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-19 = exponentials + epilogue (thread == text token == TMEM lane,
// four warps per lane quarter, 16 of the 64 sub-block columns each).  TMEM: S^T 256 columns + O 256 columns.
// =====================================================================================================================
constexpr int BT_STEP = 256;  // image tokens per S^T accumulator
constexpr int BT_SUB = 64;    // image tokens per P / V tile
constexpr int BT_PF = 2;      // L2 prefetch distance in steps
struct BtCfg {
  static constexpr int KT_BYTES = 4 * BM * BK * 2;    // resident K tile: 4 k-blocks of [128 x 64]
  static constexpr int Q_BYTES = BT_STEP * BK * 2;    // one k-block [256 n x 64 d]
  static constexpr int V_BYTES = 256 * BT_SUB * 2;    // Vv^T tile [256 d x 64 n]  /  VN: image-token tile [64 n x 256 c]
  static constexpr int P_BYTES = BM * BT_SUB * 2;     // P tile [128 t x 64 n]
  static constexpr int SMEM_BYTES = KT_BYTES + 2 * Q_BYTES + 2 * V_BYTES + 2 * P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};
static_assert(BtCfg::SMEM_BYTES <= 232448, "biattn_text: dynamic shared memory above the 227 KB per-CTA limit");
struct BtP {
  const float* stat;  // [Z][2][T]: column max, 1 / column sum   (VN: [Z][T] column max only)
  float clamp;
  int nb1, T, N, n_steps;
  int k_bc1, k_bc2, q_bc1, q_bc2, v_bc1, v_bc2;  // 1 -> batch coordinate pinned to 0
  uint32_t vn_lbo, vn_sbo;                        // VN: MN-major descriptor strides of the image-token tile
  const float* rowbias;                           // VN: optional [Z][T * rb_ld] additive score bias per text token (folded q bias)
  int rb_ld;
};

template <bool VN>
__global__ void __launch_bounds__(640, 1) biattn_text_kernel(const __grid_constant__ CUtensorMap tma_k,
                                                             const __grid_constant__ CUtensorMap tma_q,
                                                             const __grid_constant__ CUtensorMap tma_v,
                                                             const __grid_constant__ CUtensorMap tma_o, const BtP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kt = smem;
  uint8_t* qring = kt + BtCfg::KT_BYTES;           // 2 x 32 KB
  uint8_t* vring = qring + 2 * BtCfg::Q_BYTES;     // 2 x 32 KB
  uint8_t* pring = vring + 2 * BtCfg::V_BYTES;     // 2 x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(pring + 2 * BtCfg::P_BYTES);
  uint64_t* kt_full = bars;
  uint64_t* q_full = bars + 1;    // [2] Q k-block landed
  uint64_t* q_empty = bars + 3;   // [2] its four MMAs retired
  uint64_t* v_full = bars + 5;    // [2] Vv sub-block landed
  uint64_t* pv_done = bars + 7;   // [2] P_j . Vv_j retired: V slot and P slot may be overwritten
  uint64_t* p_full = bars + 9;    // [2] P tile written by the 16 exp warps
  uint64_t* s_full = bars + 11;   // S^T accumulator complete
  uint64_t* s_empty = bars + 12;  // S^T accumulator copied to registers by the 16 exp warps
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 14);
  // VN: [4 parts][128 rows] partial column sums, exchanged after the last S^T product has retired -> the K tile is free
  float* lsum_sm = reinterpret_cast<float*>(kt);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, z = blockIdx.y;
  const int z1 = z % p.nb1, z2 = z / p.nb1;
  const int NS = p.n_steps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
    mbar_init(kt_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&pv_done[s], 1);
      mbar_init(&p_full[s], 16);  // one arrival per exp warp
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 16);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_base_slot;
  const uint32_t tmem_o = tmem_s + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kt_full, BtCfg::KT_BYTES);
      for (int kb = 0; kb < 4; ++kb)
        tma_load_4d(kt + kb * (BM * BK * 2), &tma_k, kt_full, kb * BK, mt * BM, p.k_bc1 ? 0 : z1, p.k_bc2 ? 0 : z2);
      const int qz1 = p.q_bc1 ? 0 : z1, qz2 = p.q_bc2 ? 0 : z2;
      for (int qc = 0; qc < 4 * NS; ++qc) {
        const int i = qc >> 2, kb = qc & 3, s = qc & 1;
        if (i + BT_PF < NS) tma_prefetch_l2_4d(&tma_q, kb * BK, (i + BT_PF) * BT_STEP, qz1, qz2);
        if (qc >= 2) mbar_wait(&q_empty[s], ((qc >> 1) - 1) & 1);
        mbar_expect_tx(&q_full[s], BtCfg::Q_BYTES);
        tma_load_4d(qring + s * BtCfg::Q_BYTES, &tma_q, &q_full[s], kb * BK, i * BT_STEP, qz1, qz2);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      const int vz1 = p.v_bc1 ? 0 : z1, vz2 = p.v_bc2 ? 0 : z2;
      for (int sc = 0; sc < 4 * NS; ++sc) {
        const int s = sc & 1;
        if (VN) {
          // image tokens themselves, [64 n x 256 c] as four [64 n x 64 c] blocks: the B operand of P.vn is read MN-major
          if (sc + 4 * BT_PF < 4 * NS)
            for (int cb = 0; cb < 4; ++cb) tma_prefetch_l2_4d(&tma_v, cb * BK, (sc + 4 * BT_PF) * BT_SUB, vz1, vz2);
          if (sc >= 2) mbar_wait(&pv_done[s], ((sc >> 1) - 1) & 1);
          mbar_expect_tx(&v_full[s], BtCfg::V_BYTES);
          for (int cb = 0; cb < 4; ++cb)
            tma_load_4d(vring + s * BtCfg::V_BYTES + cb * (BT_SUB * 128), &tma_v, &v_full[s], cb * BK, sc * BT_SUB, vz1, vz2);
          continue;
        }
        if (sc + 4 * BT_PF < 4 * NS) tma_prefetch_l2_4d(&tma_v, (sc + 4 * BT_PF) * BT_SUB, 0, vz1, vz2);
        if (sc >= 2) mbar_wait(&pv_done[s], ((sc >> 1) - 1) & 1);
        mbar_expect_tx(&v_full[s], BtCfg::V_BYTES);
        tma_load_4d(vring + s * BtCfg::V_BYTES, &tma_v, &v_full[s], sc * BT_SUB, 0, vz1, vz2);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, 256, 0);
      mbar_wait(kt_full, 0);
      tc_fence_after();
      const uint32_t kt_addr = smem_u32(kt);
      auto issue_s = [&](int i) {  // S^T of step i: 4 k-blocks x 4 MMAs
        for (int kb = 0; kb < 4; ++kb) {
          const int qc = 4 * i + kb, s = qc & 1;
          mbar_wait(&q_full[s], (qc >> 1) & 1);
          tc_fence_after();
          const uint32_t b0 = smem_u32(qring + s * BtCfg::Q_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_s, umma_desc_k_sw128(kt_addr + kb * (BM * BK * 2) + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc,
                       (kb | k) != 0 ? 1u : 0u);
          tc_commit(&q_empty[s]);
        }
        tc_commit(s_full);
      };
      issue_s(0);
      for (int i = 0; i < NS; ++i) {
        // S^T of the NEXT step goes first: the exp warps hold step i's scores in registers (s_empty), so the tensor pipe
        // computes S^T(i+1) while they turn S^T(i) into the four P tiles, and P(i).Vv follows back to back
        if (i + 1 < NS) {
          mbar_wait(s_empty, i & 1);
          tc_fence_after();
          issue_s(i + 1);
        }
        for (int j = 0; j < 4; ++j) {
          const int sc = 4 * i + j, s = sc & 1;
          mbar_wait(&v_full[s], (sc >> 1) & 1);
          mbar_wait(&p_full[s], (sc >> 1) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(pring + s * BtCfg::P_BYTES), b0 = smem_u32(vring + s * BtCfg::V_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // VN: B = vn tile read MN-major (N = channel contiguous): 16 image tokens per MMA = 2 KB further down the block,
            // next 8 tokens (SBO) 1 KB apart, next 64 channels (LBO) one [64 n x 64 c] block = 8 KB apart
            const uint64_t db = VN ? umma_desc_mn_sw128(b0 + k * 2048, p.vn_lbo, p.vn_sbo) : umma_desc_k_sw128(b0 + k * 32);
            tc_mma_f16(tmem_o, umma_desc_k_sw128(a0 + k * 32), db, VN ? (idesc | UMMA_B_MN_MAJOR) : idesc, (sc | k) != 0 ? 1u : 0u);
          }
          tc_commit(&pv_done[s]);
        }
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, part = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const int t = mt * BM + row;
    constexpr float L2E = 1.4426950408889634f;
    float m_l2 = 0.f, inv = 0.f, lsum = 0.f, gb = 0.f;
    if (t < p.T) {
      if (VN) {
        m_l2 = -p.stat[(long)z * p.T + t] * L2E;
        if (p.rowbias) gb = p.rowbias[((long)z * p.T + t) * p.rb_ld];
      } else {
        m_l2 = -p.stat[((long)z * 2) * p.T + t] * L2E;
        inv = p.stat[((long)z * 2 + 1) * p.T + t];
      }
    }
    const float clampv = p.clamp > 0.f ? p.clamp : 3.0e38f;
    const int sw = row & 7;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const int j0 = part * 2;  // 16-byte chunks of this thread's 16 columns inside the 128-byte P row
    for (int i = 0; i < NS; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      // this thread's 16 columns of each of the four 64-token sub-blocks
      uint32_t r[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld_32x16(tmem_s + lane_addr + (uint32_t)(j * BT_SUB + part * 16), r[j]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);  // the accumulator may be overwritten by the next step's S^T
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sc = 4 * i + j, s = sc & 1;
        if (sc >= 2) mbar_wait(&pv_done[s], ((sc >> 1) - 1) & 1);  // the product that read this P slot has retired
        const int nrem = p.N - sc * BT_SUB - part * 16;  // valid columns of this thread's 16 (< 16 only at the very end)
        uint32_t h[8];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) {
          const float a = fminf(fmaxf(__uint_as_float(r[j][2 * q2]) + gb, -clampv), clampv);
          const float b = fminf(fmaxf(__uint_as_float(r[j][2 * q2 + 1]) + gb, -clampv), clampv);
          // !VN: the score as the fp16 matrix A holds it (the statistics were taken from those values);
          //  VN: the fp32 score itself (the column maxima come from the fp32 scores of the image-side kernel)
          const float2 f = VN ? make_float2(a, b) : __half22float2(__floats2half2_rn(a, b));
          float e0 = ex2_approx(fmaf(f.x, L2E, m_l2)), e1 = ex2_approx(fmaf(f.y, L2E, m_l2));
          if (nrem < 16) {  // image tokens beyond N (zero-filled q rows) must not contribute exp(-max)
            if (2 * q2 >= nrem) e0 = 0.f;
            if (2 * q2 + 1 >= nrem) e1 = 0.f;
          }
          if (VN) lsum += e0 + e1;  // the column sum of the softmax is accumulated here, no separate pass over the scores
          h[q2] = pack_half2(e0, e1);
        }
        const uint32_t prow = smem_u32(pring + s * BtCfg::P_BYTES) + row * 128;
        sts128(prow + (((j0) ^ sw) << 4), h[0], h[1], h[2], h[3]);
        sts128(prow + (((j0 + 1) ^ sw) << 4), h[4], h[5], h[6], h[7]);
        fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }
    }
    // ---- epilogue: O / colsum -> fp16 -> swizzled staging (the Q and V rings are idle by now) -> TMA store ----
    if (VN) {  // total column sum = the four column-quarter partials of this text token
      lsum_sm[part * BM + row] = lsum;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float tot = lsum_sm[row] + lsum_sm[BM + row] + lsum_sm[2 * BM + row] + lsum_sm[3 * BM + row];
      inv = tot > 0.f ? 1.f / tot : 0.f;
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const uint32_t blk = smem_u32(qring) + part * (BM * 128) + row * 128;  // this warp's 64-column block (4 x 16 KB)
    uint32_t ra[16], rb[16];
    tmem_ld_32x16(tmem_o + lane_addr + (uint32_t)(part * 64), ra);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld_wait_dep(ra);
      tmem_ld_32x16(tmem_o + lane_addr + (uint32_t)(part * 64 + c0 + 16), rb);
      uint32_t h[8];
#pragma unroll
      for (int q2 = 0; q2 < 8; ++q2) h[q2] = pack_half2(__uint_as_float(ra[2 * q2]) * inv, __uint_as_float(ra[2 * q2 + 1]) * inv);
      sts128(blk + ((((c0 >> 3)) ^ sw) << 4), h[0], h[1], h[2], h[3]);
      sts128(blk + ((((c0 >> 3) + 1) ^ sw) << 4), h[4], h[5], h[6], h[7]);
      tmem_ld_wait_dep(rb);
      if (c0 + 32 < 64) tmem_ld_32x16(tmem_o + lane_addr + (uint32_t)(part * 64 + c0 + 32), ra);
#pragma unroll
      for (int q2 = 0; q2 < 8; ++q2) h[q2] = pack_half2(__uint_as_float(rb[2 * q2]) * inv, __uint_as_float(rb[2 * q2 + 1]) * inv);
      sts128(blk + ((((c0 >> 3) + 2) ^ sw) << 4), h[0], h[1], h[2], h[3]);
      sts128(blk + ((((c0 >> 3) + 3) ^ sw) << 4), h[4], h[5], h[6], h[7]);
    }
    tc_fence_before();
    fence_proxy_async();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (warp == 4 && lane == 0) {
      for (int cb = 0; cb < 4; ++cb) tma_store_4d(&tma_o, qring + cb * (BM * 128), cb * 64, mt * BM, z1, z2);
      tma_store_commit_and_wait_read();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_s, 512);
  }
}

// =====================================================================================================================
// Image -> text side, fused with the out-projection, layer scale and residual.
//   vn [B][N][256] fp16 = the layer-normed image tokens
//   gT [B][H][T][256 c] fp16 = K_h (d^-1/2 Wv_h): the query projection of head h folded into the key operand —
//                              S = (vn Wv_h^T + bv) d^-1/2 K_h^T == vn gT_h^T + gbias_h[t] — so q is never built and ONE image-token
//                              tile serves all heads;  gbias [B*H][T] fp32 = d^-1/2 bv_h . K_h[t]
//   mT [B][H][256 o][T] fp16 = (W_h V_l,h^T): the value AND output projections of head h folded into one [T x 256] operand per
//                              (image, head) by a tiny GEMM — (P V_l,h) W_h^T == P (V_l,h W_h^T), so the per-head context O never
//                              exists and a head costs two tensor-core products instead of three
//   out[B][N][256] fp16
// Persistent: one CTA per SM walks (image, 128-token tile) items; per item the eight heads run back to back:
//   S (128 x 256 t)   = vn . gT_h^T (+ gbias_h)   16 MMAs 128x256x16, the operands streamed as [.. x 64] k-blocks through a 3-stage ring
//   P                 = softmax_t(clamp(S) + mask): thread == image token == TMEM lane, 4 warps per lane quarter, 64 tokens each;
//                       pass 1 over the TMEM columns = column maxima (for the text side) + partial row maxima, pass 2 = exp into
//                       registers + partial row sums; both statistics are exchanged through shared memory; P = e / rowsum -> fp16
//                       -> shared memory (A operand)
//   D (128 x 256 o)  += P . mT_h^T           second TMEM accumulator, summed over the heads
//   out               = res + gamma * (D + bias)
// The scratch columns are free as soon as pass 2 has read them, so the issuer runs S of head h+1 while the warps normalise and
// store P of head h, and D of head h while they exponentiate S of head h+1.
// TMEM: columns 0..255 S, 256..511 D.  Shared memory: one 64 KB tile PX (P, finally the output staging), 3 x 48 KB ring stages
// ([128 x 64] A slot + [256 x 64] B slot), exchange arrays.
// Column maxima of the clamped scores over the tile's valid rows go to colmax_part[b][h][tile][t] (reduced by
// colmax_reduce_kernel); the text side needs them as the softmax shift.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-19 = softmax / epilogue.
// =====================================================================================================================
constexpr int BI_STAGES = 3;
struct BiCfg {
  static constexpr int PX_BYTES = 4 * BM * BK * 2;           // 64 KB: four k-blocks [128 x 64]
  static constexpr int A_BYTES = BM * BK * 2;                // 16 KB
  static constexpr int B_BYTES = 256 * BK * 2;               // 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 48 KB
  static constexpr int MAX_HEADS = 8;
  static constexpr int EXTRA_BYTES = (4 * BM * 2 + 4 * 256 + MAX_HEADS * 256) * 4;  // rx | rsum, cmx, gk
  static constexpr int SMEM_BYTES = PX_BYTES + BI_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EXTRA_BYTES;
};
static_assert(BiCfg::SMEM_BYTES <= 232448, "biattn_image: dynamic shared memory above the 227 KB per-CTA limit");
struct BiP {
  const float* mask;   // [B][T] 1 keep / 0 padding, or nullptr
  const float* bias;   // [256] out-projection bias
  const float* gamma;  // [256] layer scale, or nullptr (1)
  const __half* res;   // residual [B][N][256] (row stride res_ld, batch stride res_b) or nullptr
  long res_ld, res_b;
  float* colmax_part;  // [B][H][tiles_per_img][T]
  const float* gbias;  // [B*H][T * gb_ld] additive score bias per (head, token): the folded query-projection bias, or nullptr
  int gb_ld;
  float clamp;
  int B, H, T, N, tiles_per_img, total_tiles;
};

__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(640, 1) biattn_image_kernel(const __grid_constant__ CUtensorMap tma_q,
                                                              const __grid_constant__ CUtensorMap tma_k,
                                                              const __grid_constant__ CUtensorMap tma_m,
                                                              const __grid_constant__ CUtensorMap tma_o, const BiP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* px = smem;
  uint8_t* ring = px + BiCfg::PX_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + BI_STAGES * BiCfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [3] ring stage landed
  uint64_t* empty_bar = bars + BI_STAGES;        // [3] its MMAs retired
  uint64_t* s_full = bars + 6;                   // S complete
  uint64_t* scratch_free = bars + 7;             // S has been read (16 warps): the scratch columns may take the next head's S
  uint64_t* p_ready = bars + 8;                  // P written (16 warps)
  uint64_t* op_done = bars + 9;                  // the head's P . mT MMAs retired: PX may be rewritten; last head: D complete
  uint64_t* dv_free = bars + 10;                 // D has been read (16 warps)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* rx = reinterpret_cast<float*>(bars + 32);  // [4 parts][128 rows] partial row maxima
  float* rsum = rx + 4 * BM;                         // [4 parts][128 rows] partial row sums
  float* cmx = rsum + 4 * BM;                        // [4 lane quarters][256 tokens] column maxima
  float* gk = cmx + 4 * 256;                         // [H][256] additive row-softmax term per (head, token): folded query bias
                                                     // + token mask (0 keep / -inf masked)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_m);
    tma_prefetch_desc(&tma_o);
    for (int s = 0; s < BI_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(scratch_free, 16);
    mbar_init(p_ready, 16);
    mbar_init(op_done, 1);
    mbar_init(dv_free, 16);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_base_slot;
  const uint32_t tmem_d = tmem_s + 256;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_img, m0 = (tile - b * p.tiles_per_img) * BM;
        // consumption order of the MMA issuer: S(0); then per head  S(h+1), D(h)
        auto load = [&](int ph, int h) {
          for (int kb = 0; kb < 4; ++kb, ++it) {
            const int s = it % BI_STAGES;
            mbar_wait(&empty_bar[s], ((it / BI_STAGES) & 1) ^ 1);
            uint8_t* st = ring + s * BiCfg::STAGE_BYTES;
            if (ph == 0) {  // S: A = image-token k-block (the same for every head: L2 hit), B = gT_h k-block
              mbar_expect_tx(&full_bar[s], BiCfg::STAGE_BYTES);
              tma_load_4d(st, &tma_q, &full_bar[s], kb * BK, m0, b, 0);
              tma_load_4d(st + BiCfg::A_BYTES, &tma_k, &full_bar[s], kb * BK, 0, h, b);
            } else {  // D: B = mT_h [256 o x 64 t]
              mbar_expect_tx(&full_bar[s], BiCfg::B_BYTES);
              tma_load_4d(st + BiCfg::A_BYTES, &tma_m, &full_bar[s], kb * BK, 0, h, b);
            }
          }
        };
        load(0, 0);
        for (int h = 0; h < H; ++h) {
          if (h + 1 < H) load(0, h + 1);
          load(1, h);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, 256, 0);
      int it = 0, hc = 0, tcnt = 0;
      const uint32_t px_addr = smem_u32(px);
      auto issue = [&](int ph, int h) {
        // ph 0: S = Q_h K_h^T -> scratch columns;  1: D += P mT_h^T
        for (int kb = 0; kb < 4; ++kb, ++it) {
          const int s = it % BI_STAGES;
          mbar_wait(&full_bar[s], (it / BI_STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(ring + s * BiCfg::STAGE_BYTES);
          const uint32_t a0 = ph == 0 ? st : px_addr + kb * BiCfg::A_BYTES, b0 = st + BiCfg::A_BYTES;
          const uint32_t acc = ph == 0 ? tmem_s : tmem_d;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(acc, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc,
                       ((ph == 1 ? h : 0) | kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(ph == 0 ? s_full : op_done);
      };
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcnt) {
        // S of the item's first head (the previous item's last S must have been read out of the scratch columns)
        if (hc > 0) {
          mbar_wait(scratch_free, (hc - 1) & 1);
          tc_fence_after();
        }
        issue(0, 0);
        for (int h = 0; h < H; ++h, ++hc) {
          if (h + 1 < H) {  // the next head's S only needs the scratch columns: it runs while P of this head is written
            mbar_wait(scratch_free, hc & 1);
            tc_fence_after();
            issue(0, h + 1);
          }
          mbar_wait(p_ready, hc & 1);
          tc_fence_after();
          if (h == 0 && tcnt > 0) {  // the previous item's D must have been read by the epilogue
            mbar_wait(dv_free, (tcnt - 1) & 1);
            tc_fence_after();
          }
          issue(1, h);
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, part = (warp - 4) >> 2;
    const int row = ew * 32 + lane, tid_e = threadIdx.x - 128;
    const int sw = row & 7;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t px_row = smem_u32(px) + part * BiCfg::A_BYTES + row * 128;  // this thread's 128-byte row of k-block `part`
    const uint32_t gk_addr0 = smem_u32(gk) + part * 64 * 4, rx_addr = smem_u32(rx) + row * 4, rsum_addr = smem_u32(rsum) + row * 4;
    constexpr float L2E = 1.4426950408889634f;
    const float NEG_INF = __int_as_float(0xff800000);
    const float clampv = p.clamp > 0.f ? p.clamp : 3.0e38f;
    const bool issuer = (tid_e == 0);
    int hc = 0, tcnt = 0, cur_b = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcnt) {
      const int b = tile / p.tiles_per_img, ti = tile - b * p.tiles_per_img, m0 = ti * BM;
      const bool valid_row = m0 + row < p.N;
      const bool full_tile = m0 + BM <= p.N;
      if (b != cur_b) {  // token mask of this image (warp-uniform branch: every thread walks the same items)
        asm volatile("bar.sync 1, 512;" ::: "memory");  // nobody still reads the previous image's flags
        if (tid_e < 256) {
          const bool kept = tid_e < p.T && (!p.mask || p.mask[(long)b * p.T + tid_e] != 0.f);
          for (int hh = 0; hh < H; ++hh)
            gk[hh * 256 + tid_e] = !kept ? NEG_INF : (p.gbias ? p.gbias[(((long)b * H + hh) * p.T + tid_e) * p.gb_ld] : 0.f);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        cur_b = b;
      }
      for (int h = 0; h < H; ++h, ++hc) {
        // ================= S -> P = softmax_t(clamp(S) + mask) =================
        mbar_wait(s_full, hc & 1);
        tc_fence_after();
        const uint32_t t_s = tmem_s + lane_addr + (uint32_t)(part * 64);
        const uint32_t kl_addr = gk_addr0 + (uint32_t)(h * 256 * 4);  // this head's additive term (query bias + token mask)
        // ---- pass 1 over this thread's 64 tokens: column maxima (for the text side) and the partial row maximum ----
        // The +-5e4 clamp of the scores (fuse_helper.py:245-252) is applied to the REDUCED values (clamp is monotonic:
        // max(clamp(s)) == clamp(max(s))), not per element; `kl` carries the token mask as an additive 0 / -inf.
        float mpart = NEG_INF;
        {
          uint32_t ra[16], rb[16];
          tmem_ld_32x16(t_s, ra);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t (&r)[16] = (j & 1) ? rb : ra;
            tmem_ld_wait_dep(r);
            if (j + 1 < 4) tmem_ld_32x16(t_s + (uint32_t)((j + 1) * 16), (j & 1) ? ra : rb);
            float cm[16];
            if (full_tile) {  // warp-uniform: every row of the tile is a real image token
#pragma unroll
              for (int i = 0; i < 16; ++i) cm[i] = warp_redux_max(__uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) cm[i] = warp_redux_max(valid_row ? __uint_as_float(r[i]) : NEG_INF);
            }
            if (lane == 0) {  // the reduction result is warp-uniform: one lane stores the 16 column maxima of this warp's 32 rows
              const uint32_t dst = smem_u32(cmx) + (uint32_t)((ew * 256 + part * 64 + j * 16) * 4);
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4)
                sts128(dst + i4 * 16, __float_as_uint(cm[i4 * 4]), __float_as_uint(cm[i4 * 4 + 1]), __float_as_uint(cm[i4 * 4 + 2]),
                       __float_as_uint(cm[i4 * 4 + 3]));
            }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 kf = lds128f(kl_addr + (uint32_t)((j * 16 + i4 * 4) * 4));
              mpart = fmaxf(mpart, __uint_as_float(r[i4 * 4 + 0]) + kf.x);
              mpart = fmaxf(mpart, __uint_as_float(r[i4 * 4 + 1]) + kf.y);
              mpart = fmaxf(mpart, __uint_as_float(r[i4 * 4 + 2]) + kf.z);
              mpart = fmaxf(mpart, __uint_as_float(r[i4 * 4 + 3]) + kf.w);
            }
          }
        }
        rx[part * BM + row] = mpart;
        // the previous item's output staging (PX) must have been read by the TMA store before P overwrites it
        if (issuer && h == 0 && tcnt > 0) tma_store_wait_read_all();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (tid_e < p.T) {  // column-max partial of this (image, head, tile), clamped like the scores
          float c4 = fmaxf(fmaxf(cmx[tid_e], cmx[256 + tid_e]), fmaxf(cmx[512 + tid_e], cmx[768 + tid_e]));
          if (p.gbias) c4 += p.gbias[(((long)b * H + h) * p.T + tid_e) * p.gb_ld];  // the text side sees the full score, no mask
          c4 = fminf(fmaxf(c4, -clampv), clampv);
          p.colmax_part[(((long)b * H + h) * p.tiles_per_img + ti) * p.T + tid_e] = c4;
        }
        const float m_raw = fmaxf(fmaxf(lds32f(rx_addr), lds32f(rx_addr + BM * 4)), fmaxf(lds32f(rx_addr + 2 * BM * 4), lds32f(rx_addr + 3 * BM * 4)));
        // every token masked: the reference's fp32 sum A + (-9e15) swallows A and the softmax is uniform over the T tokens
        const bool uniform = (m_raw == NEG_INF);
        const float m = fminf(fmaxf(m_raw, -clampv), clampv);  // max over the kept tokens of the CLAMPED scores
        const float ml2 = uniform ? 0.f : m * L2E;
        // per-element clamping can only change exp(clamp(s) - m) when the row maximum itself exceeds the clamp (s > 5e4), or
        // when m is so low that exp(-5e4 - m) does not underflow; otherwise exp(s - m) is bit-identical without it
        const bool need_clamp = (m_raw > clampv) || (m_raw < -clampv + 256.f);
        // ---- pass 2: e = exp(S - rowmax) kept in registers, partial row sum ----
        float e[64];
        float lpart = 0.f;
        {
          uint32_t ra[16], rb[16];
          tmem_ld_32x16(t_s, ra);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t (&r)[16] = (j & 1) ? rb : ra;
            tmem_ld_wait_dep(r);
            if (j + 1 < 4) tmem_ld_32x16(t_s + (uint32_t)((j + 1) * 16), (j & 1) ? ra : rb);
            if (j == 3) {  // S has left the scratch columns: they may take the next head's S
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(scratch_free);
            }
            if (!need_clamp && !uniform) {  // the common path: 4 issue slots per score (FADD, FFMA, MUFU.EX2, FADD)
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 kf = lds128f(kl_addr + (uint32_t)((j * 16 + i4 * 4) * 4));
                const float kk[4] = {kf.x, kf.y, kf.z, kf.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // kk = folded query bias of this (head, token), or -inf on a masked token: exp2(-inf) = 0, exactly what the
                  // reference's -9e15 does in fp32
                  const float ev = ex2_approx(fmaf(__uint_as_float(r[i4 * 4 + i]) + kk[i], L2E, -ml2));
                  e[j * 16 + i4 * 4 + i] = ev;
                  lpart += ev;
                }
              }
            } else {
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 kf = lds128f(kl_addr + (uint32_t)((j * 16 + i4 * 4) * 4));
                const float kk[4] = {kf.x, kf.y, kf.z, kf.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int c = j * 16 + i4 * 4 + i;
                  const float sc = fminf(fmaxf(__uint_as_float(r[i4 * 4 + i]) + kk[i], -clampv), clampv);
                  float ev = (kk[i] == NEG_INF) ? 0.f : ex2_approx(fmaf(sc, L2E, -ml2));
                  if (uniform) ev = (part * 64 + c < p.T) ? 1.f : 0.f;
                  e[c] = ev;
                  lpart += ev;
                }
              }
            }
          }
        }
        rsum[part * BM + row] = lpart;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float inv = 1.f / (lds32f(rsum_addr) + lds32f(rsum_addr + BM * 4) + lds32f(rsum_addr + 2 * BM * 4) + lds32f(rsum_addr + 3 * BM * 4));
        if (hc > 0) mbar_wait(op_done, (hc - 1) & 1);  // the previous head's product no longer reads PX
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(px_row + ((j ^ sw) << 4), pack_half2(e[8 * j] * inv, e[8 * j + 1] * inv), pack_half2(e[8 * j + 2] * inv, e[8 * j + 3] * inv),
                 pack_half2(e[8 * j + 4] * inv, e[8 * j + 5] * inv), pack_half2(e[8 * j + 6] * inv, e[8 * j + 7] * inv));
        fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
      }
      // ================= item epilogue: out = res + gamma * (D + bias) =================
      mbar_wait(op_done, (hc - 1) & 1);  // the last head's product: D complete, PX free
      tc_fence_after();
      {
        const bool has_res = p.res != nullptr && valid_row;
        const uint4* src = reinterpret_cast<const uint4*>(p.res + (long)b * p.res_b + (long)(m0 + row) * p.res_ld + part * 64);
        const uint32_t t_d = tmem_d + lane_addr + (uint32_t)(part * 64);
        const float4* gam4 = reinterpret_cast<const float4*>(p.gamma) + part * 16;
        const float4* bia4 = reinterpret_cast<const float4*>(p.bias) + part * 16;
        uint32_t ra[16], rb[16];
        tmem_ld_32x16(t_d, ra);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 16 columns = two 16-byte chunks per step
          uint32_t (&r)[16] = (j & 1) ? rb : ra;
          uint4 rq0 = make_uint4(0, 0, 0, 0), rq1 = make_uint4(0, 0, 0, 0);
          if (has_res) {
            rq0 = __ldg(src + 2 * j);
            rq1 = __ldg(src + 2 * j + 1);
          }
          tmem_ld_wait_dep(r);
          if (j + 1 < 4) tmem_ld_32x16(t_d + (uint32_t)((j + 1) * 16), (j & 1) ? ra : rb);
          if (j == 3) {  // D has been read: the next item's products may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dv_free);
          }
          float o[16];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 sv = p.gamma ? __ldg(gam4 + j * 4 + i4) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 bv = p.bias ? __ldg(bia4 + j * 4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            o[i4 * 4 + 0] = (__uint_as_float(r[i4 * 4 + 0]) + bv.x) * sv.x;  // gamma * (D + bias)
            o[i4 * 4 + 1] = (__uint_as_float(r[i4 * 4 + 1]) + bv.y) * sv.y;
            o[i4 * 4 + 2] = (__uint_as_float(r[i4 * 4 + 2]) + bv.z) * sv.z;
            o[i4 * 4 + 3] = (__uint_as_float(r[i4 * 4 + 3]) + bv.w) * sv.w;
          }
          const __half2* h0 = reinterpret_cast<const __half2*>(&rq0);
          const __half2* h1 = reinterpret_cast<const __half2*>(&rq1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f0 = __half22float2(h0[i]), f1 = __half22float2(h1[i]);
            o[2 * i] += f0.x;
            o[2 * i + 1] += f0.y;
            o[8 + 2 * i] += f1.x;
            o[8 + 2 * i + 1] += f1.y;
          }
          sts128(px_row + (((2 * j) ^ sw) << 4), pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]),
                 pack_half2(o[6], o[7]));
          sts128(px_row + (((2 * j + 1) ^ sw) << 4), pack_half2(o[8], o[9]), pack_half2(o[10], o[11]), pack_half2(o[12], o[13]),
                 pack_half2(o[14], o[15]));
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (issuer) {
        for (int cb = 0; cb < 4; ++cb) tma_store_4d(&tma_o, px + cb * BiCfg::A_BYTES, cb * 64, m0, b, 0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_read_all();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_s, 512);
  }
}

// colmax[z][t] = max over the tiles of colmax_part[z][tile][t]
__global__ void colmax_reduce_kernel(const float* __restrict__ part, int tiles, int T, float* __restrict__ out) {
  const int z = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float m = __int_as_float(0xff800000);
    const float* src = part + (long)z * tiles * T + t;
    for (int i = 0; i < tiles; ++i) m = fmaxf(m, src[(long)i * T]);
    out[(long)z * T + t] = m;
  }
}

}  // namespace mqdet

using namespace mqdet;

static int text_launch(bool vn, const float* rowbias, int64_t rb_ld, const void* k, int64_t k_ld, int64_t k_b1, int64_t k_b2, const void* q, int64_t q_ld, int64_t q_b1,
                       int64_t q_b2, const void* v, int64_t v_ld, int64_t v_b1, int64_t v_b2, const float* stat, float clamp,
                       void* out, int64_t o_ld, int64_t o_b1, int64_t o_b2, int64_t nb1, int64_t nb2, int64_t T, int64_t N,
                       int64_t Np, int64_t d, void* stream) {
  MQ_REQUIRE(k && q && v && stat && out, "biattn_text: null pointer");
  MQ_REQUIRE(d == 256, "biattn_text: head dim must be 256 (embed 2048 / 8 heads, fuse_helper.py:186-189), got %ld", (long)d);
  MQ_REQUIRE(T >= 1 && T <= 256 && N >= 1 && Np >= N && (Np % 8) == 0, "biattn_text: need 1 <= T <= 256, Np >= N, Np %% 8 == 0");
  MQ_REQUIRE(nb1 >= 1 && nb2 >= 1 && nb1 * nb2 <= 65535, "biattn_text: bad batch");
  const int64_t lds[] = {k_ld, k_b1, k_b2, q_ld, q_b1, q_b2, v_ld, v_b1, v_b2, o_ld, o_b1, o_b2};
  for (int64_t x : lds) MQ_REQUIRE((x % 8) == 0, "biattn_text: strides must be multiples of 8 elements");
  MQ_REQUIRE(((uintptr_t)k % 16) == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "biattn_text: operands must be 16-byte aligned");
  CUtensorMap mk, mq, mv, mo;
  BtP p;
  memset(&p, 0, sizeof(p));
  int rc = make_operand_map(&mk, k, T, d, k_ld, (int)nb1, k_b1, (int)nb2, k_b2, BM, &p.k_bc1, &p.k_bc2);
  if (rc) return rc;
  rc = make_operand_map(&mq, q, N, d, q_ld, (int)nb1, q_b1, (int)nb2, q_b2, BT_STEP, &p.q_bc1, &p.q_bc2);
  if (rc) return rc;
  if (vn)  // image tokens [N][256 c]: box [64 n x 64 c]
    rc = make_operand_map(&mv, v, N, 256, v_ld, (int)nb1, v_b1, (int)nb2, v_b2, BT_SUB, &p.v_bc1, &p.v_bc2);
  else     // Vv^T [256 d][Np]: box [256 d x 64 n]
    rc = make_operand_map(&mv, v, d, Np, v_ld, (int)nb1, v_b1, (int)nb2, v_b2, 256, &p.v_bc1, &p.v_bc2);
  if (rc) return rc;
  MQ_REQUIRE((nb1 == 1 || o_b1 != 0) && (nb2 == 1 || o_b2 != 0), "biattn_text: output batch strides must be non-zero");
  rc = make_store_map(&mo, out, MQDET_F16, T, d, o_ld, (int)nb1, o_b1, (int)nb2, o_b2);
  if (rc) return rc;
  p.stat = stat;
  p.clamp = clamp;
  p.nb1 = (int)nb1;
  p.T = (int)T;
  p.N = (int)N;
  p.n_steps = (int)((N + BT_STEP - 1) / BT_STEP);
  p.rowbias = rowbias;
  p.rb_ld = (int)(rb_ld > 0 ? rb_ld : 1);
  p.vn_lbo = BT_SUB * 128;  // next 64 channels: the next [64 n x 64 c] block
  p.vn_sbo = 1024;          // next 8 image tokens
  if (const char* dbg = getenv("MQDET_VN_DESC_SWAP")) {  // bring-up aid: swap the two strides
    if (atoi(dbg)) { p.vn_lbo = 1024; p.vn_sbo = BT_SUB * 128; }
  }
  dim3 grid((unsigned)((T + BM - 1) / BM), (unsigned)(nb1 * nb2));
  if (vn) {
    rc = ensure_dyn_smem(reinterpret_cast<const void*>(&biattn_text_kernel<true>), BtCfg::SMEM_BYTES);
    if (rc) return rc;
    biattn_text_kernel<true><<<grid, 640, BtCfg::SMEM_BYTES, (cudaStream_t)stream>>>(mk, mq, mv, mo, p);
  } else {
    rc = ensure_dyn_smem(reinterpret_cast<const void*>(&biattn_text_kernel<false>), BtCfg::SMEM_BYTES);
    if (rc) return rc;
    biattn_text_kernel<false><<<grid, 640, BtCfg::SMEM_BYTES, (cudaStream_t)stream>>>(mk, mq, mv, mo, p);
  }
  return check_launch("biattn_text_kernel");
}

extern "C" int mqdet_biattn_text(const void* k, int64_t k_ld, int64_t k_b1, int64_t k_b2, const void* q, int64_t q_ld,
                                 int64_t q_b1, int64_t q_b2, const void* vvT, int64_t v_ld, int64_t v_b1, int64_t v_b2,
                                 const float* stat, float clamp, void* out, int64_t o_ld, int64_t o_b1, int64_t o_b2,
                                 int64_t nb1, int64_t nb2, int64_t T, int64_t N, int64_t Np, int64_t d, void* stream) {
  return text_launch(false, nullptr, 0, k, k_ld, k_b1, k_b2, q, q_ld, q_b1, q_b2, vvT, v_ld, v_b1, v_b2, stat, clamp, out, o_ld, o_b1, o_b2,
                     nb1, nb2, T, N, Np, d, stream);
}

extern "C" int mqdet_biattn_text_vn(const void* k, int64_t k_ld, int64_t k_b1, int64_t k_b2, const void* q, int64_t q_ld,
                                    int64_t q_b1, int64_t q_b2, const void* vn, int64_t vn_ld, int64_t vn_b1, int64_t vn_b2,
                                    const float* colmax, const float* rowbias, int64_t rb_ld, float clamp, void* out, int64_t o_ld,
                                    int64_t o_b1, int64_t o_b2, int64_t nb1, int64_t nb2, int64_t T, int64_t N, void* stream) {
  return text_launch(true, rowbias, rb_ld, k, k_ld, k_b1, k_b2, q, q_ld, q_b1, q_b2, vn, vn_ld, vn_b1, vn_b2, colmax, clamp, out, o_ld, o_b1,
                     o_b2, nb1, nb2, T, N, (N + 7) / 8 * 8, 256, stream);
}

extern "C" int64_t mqdet_biattn_image_workspace_floats(int64_t B, int64_t H, int64_t N, int64_t T) {
  return B * H * ((N + BM - 1) / BM) * T;
}

extern "C" int mqdet_biattn_image(const void* vn, int64_t vn_ld, int64_t vn_b, const void* gT, int64_t g_ld, int64_t g_bh,
                                  int64_t g_b, const float* gbias, int64_t gb_ld, const void* mT, int64_t m_ld, int64_t m_bh,
                                  int64_t m_b, const float* bias, const float* gamma, const void* res, int64_t res_ld,
                                  int64_t res_b, const float* mask, float clamp, void* out, int64_t o_ld, int64_t o_b,
                                  float* colmax, float* workspace, int64_t B, int64_t H, int64_t N, int64_t T, void* stream) {
  MQ_REQUIRE(vn && gT && mT && out && colmax && workspace, "biattn_image: null pointer");
  MQ_REQUIRE(B >= 1 && H >= 1 && H <= BiCfg::MAX_HEADS && N >= 1 && T >= 8 && T <= 256 && (T % 8) == 0,
             "biattn_image: need B, N >= 1, 1 <= H <= 8, 8 <= T <= 256, T %% 8 == 0 (got B=%ld H=%ld N=%ld T=%ld)", (long)B, (long)H, (long)N, (long)T);
  const int64_t lds[] = {vn_ld, vn_b, g_ld, g_bh, g_b, m_ld, m_bh, m_b, o_ld, o_b, res ? res_ld : 0, res ? res_b : 0};
  for (int64_t x : lds) MQ_REQUIRE((x % 8) == 0, "biattn_image: strides must be multiples of 8 elements");
  MQ_REQUIRE(((uintptr_t)vn % 16) == 0 && ((uintptr_t)gT % 16) == 0 && ((uintptr_t)mT % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
                 ((uintptr_t)res % 16) == 0 && ((uintptr_t)bias % 16) == 0 && ((uintptr_t)gamma % 16) == 0,
             "biattn_image: operands must be 16-byte aligned");
  CUtensorMap mq, mk, mm, mo;
  int bc1, bc2;
  int rc = make_operand_map(&mq, vn, N, 256, vn_ld, (int)B, vn_b, 1, 0, BM, &bc1, &bc2);
  if (rc) return rc;
  MQ_REQUIRE((H == 1 || (g_bh != 0 && m_bh != 0)) && (B == 1 || (g_b != 0 && m_b != 0)), "biattn_image: head / image strides must be non-zero");
  rc = make_operand_map(&mk, gT, T, 256, g_ld, (int)H, g_bh, (int)B, g_b, 256, &bc1, &bc2);  // [B][H][T][256 c]
  if (rc) return rc;
  rc = make_operand_map(&mm, mT, 256, T, m_ld, (int)H, m_bh, (int)B, m_b, 256, &bc1, &bc2);  // [B][H][256 o][T]
  if (rc) return rc;
  MQ_REQUIRE(B == 1 || o_b != 0, "biattn_image: output batch stride must be non-zero");
  rc = make_store_map(&mo, out, MQDET_F16, N, 256, o_ld, (int)B, o_b, 1, 0);
  if (rc) return rc;
  BiP p;
  memset(&p, 0, sizeof(p));
  p.mask = mask;
  p.bias = bias;
  p.gamma = gamma;
  p.res = (const __half*)res;
  p.res_ld = res_ld;
  p.res_b = res_b;
  p.colmax_part = workspace;
  p.gbias = gbias;
  p.gb_ld = (int)(gb_ld > 0 ? gb_ld : 1);
  p.clamp = clamp;
  p.B = (int)B; p.H = (int)H; p.T = (int)T; p.N = (int)N;
  p.tiles_per_img = (int)((N + BM - 1) / BM);
  p.total_tiles = p.tiles_per_img * (int)B;
  rc = ensure_dyn_smem(reinterpret_cast<const void*>(&biattn_image_kernel), BiCfg::SMEM_BYTES);
  if (rc) return rc;
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  biattn_image_kernel<<<grid, 640, BiCfg::SMEM_BYTES, (cudaStream_t)stream>>>(mq, mk, mm, mo, p);
  rc = check_launch("biattn_image_kernel");
  if (rc) return rc;
  colmax_reduce_kernel<<<(unsigned)(B * H), 256, 0, (cudaStream_t)stream>>>(workspace, p.tiles_per_img, (int)T, colmax);
  return check_launch("colmax_reduce_kernel");
}
