// EXPERIMENTAL — not built into libmqdet_b200.so, never executed on a GPU so far (see experimental/README.md).
//
// gemm_plain16_kernel<STAGES, BRES>: the persistent 128x256 tcgen05 GEMM of mqdet_b200/csrc/gemm.cu with SIXTEEN epilogue
// warps for the "plain" products (fp16 output through the TMA store, v = acc * S + T with S / T uniform or per column, optional
// clamp; no activation, no residual).  640 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3
// idle, warps 4-19 epilogue: four warps per TMEM lane quarter, each converting 16 of the 64 columns of a staging window, so
// every scheduler has four epilogue warps to hide the tcgen05.ld / LDS / F2FP latencies (eight warps leave two per scheduler).
// Producer, MMA issue, work distribution and barriers are those of gemm_tcp_kernel; tmem_empty expects 16 arrivals.
#include "../mqdet_b200/csrc/gemm.cu"

namespace mqdet {

template <int STAGES, bool BRES>
__global__ void __launch_bounds__(640, 1) gemm_plain16_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                              const __grid_constant__ CUtensorMap tma_b,
                                                              const __grid_constant__ CUtensorMap tma_c, const GemmP p,
                                                              int tiles_m, int tiles_n, int total_items, int mc) {
  constexpr int BN = 256;
  using Cfg = TcpCfg<BN, STAGES, BRES>;
  static_assert(Cfg::STG_BYTES >= 2 * BM * 128, "two 64-column staging windows");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* stg = smem + Cfg::RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::STG_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint64_t* b_full_bar = bars + 2 * STAGES + 4;
  uint64_t* b_empty_bar = bars + 2 * STAGES + 5;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  float* s_vec = reinterpret_cast<float*>(bars + 32);
  float* t_vec = s_vec + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (int)((p.K + BK - 1) / BK);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 16);  // one arrival per epilogue warp
    }
    mbar_init(b_full_bar, 1);
    mbar_init(b_empty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // work distribution: identical to gemm_tcp_kernel
  const int share = BRES ? (total_items + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int w_begin = BRES ? (int)blockIdx.x * share : (int)blockIdx.x;
  const int w_end = BRES ? min(total_items, w_begin + share) : total_items;
  auto next_item = [&](int& cur, int& z, int& m0, int& mcount, int& n_tile) -> bool {
    if (cur >= w_end) return false;
    if (BRES) {
      const int per_z = tiles_m * tiles_n;
      z = cur / per_z;
      const int u = cur - z * per_z;
      const int c = u / (mc * tiles_n);
      const int v = u - c * mc * tiles_n;
      const int sc = min(mc, tiles_m - c * mc);
      n_tile = v / sc;
      const int mi = v - n_tile * sc;
      m0 = c * mc + mi;
      mcount = min(sc - mi, w_end - cur);
      cur += mcount;
    } else {
      n_tile = cur % tiles_n;
      m0 = (cur / tiles_n) % tiles_m;
      z = cur / (tiles_n * tiles_m);
      mcount = 1;
      cur += gridDim.x;
    }
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, li = 0;
      int z, m0, mcount, n_tile;
      for (int cur = w_begin; next_item(cur, z, m0, mcount, n_tile); ++li) {
        const int z1 = z % p.nb1, z2 = z / p.nb1;
        const int az1 = p.a_bcast1 ? 0 : z1, az2 = p.a_bcast2 ? 0 : z2;
        const int bz1 = p.b_bcast1 ? 0 : z1, bz2 = p.b_bcast2 ? 0 : z2;
        if (BRES) {
          mbar_wait(b_empty_bar, (li & 1) ^ 1);
          mbar_expect_tx(b_full_bar, num_kb * Cfg::B_BYTES);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_4d(smem_b + kb * Cfg::B_BYTES, &tma_b, b_full_bar, kb * BK, n_tile * BN, bz1, bz2);
        }
        for (int mt = 0; mt < mcount; ++mt)
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
            tma_load_4d(smem_a + s * Cfg::A_BYTES, &tma_a, &full_bar[s], kb * BK, (m0 + mt) * BM, az1, az2);
            if (!BRES) tma_load_4d(smem_b + s * Cfg::B_BYTES, &tma_b, &full_bar[s], kb * BK, n_tile * BN, bz1, bz2);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0);
      int it = 0, lt = 0, li = 0;
      int z, m0, mcount, n_tile;
      for (int cur = w_begin; next_item(cur, z, m0, mcount, n_tile); ++li) {
        if (BRES) {
          mbar_wait(b_full_bar, li & 1);
          tc_fence_after();
        }
        for (int mt = 0; mt < mcount; ++mt, ++lt) {
          const int buf = lt & 1;
          mbar_wait(&tmem_empty_bar[buf], ((lt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + s * Cfg::A_BYTES);
            const uint32_t b_addr = smem_u32(smem_b + (BRES ? kb : s) * Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_f16(acc, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[s]);
          }
          tc_commit(&tmem_full_bar[buf]);
        }
        if (BRES) tc_commit(b_empty_bar);
      }
    }
  } else if (warp >= 4) {
    // ---- 16 epilogue warps ----
    const int ew = (warp - 4) & 3, part = (warp - 4) >> 2, tid_e = threadIdx.x - 128;
    const int row = ew * 32 + lane, sw = row & 7;
    const bool issuer = (warp == 4 && lane == 0);
    const bool bcol = p.bias_mode == MQDET_VEC_PER_COL, brow = p.bias_mode == MQDET_VEC_PER_ROW;
    const bool gcol = p.gate_mode == MQDET_VEC_PER_COL, grow = p.gate_mode == MQDET_VEC_PER_ROW;
    const bool vec = bcol || gcol;
    const float bscale = p.scale_after_bias ? p.alpha : 1.f;
    float g_u = 1.f;
    if (p.gate_mode == MQDET_VEC_SCALAR) g_u = p.gate_tanh ? tanhf(p.gate[0]) : p.gate[0];
    const float clampv = p.clamp;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t s_addr = smem_u32(s_vec), t_addr = smem_u32(t_vec);
    int lt = 0, wcount = 0;
    int z, m0, mcount, n_tile;
    for (int cur = w_begin; next_item(cur, z, m0, mcount, n_tile);) {
      const int z1 = z % p.nb1, z2 = z / p.nb1;
      for (int mt = 0; mt < mcount; ++mt, ++lt) {
        const int buf = lt & 1;
        // per-column / per-row parameters are fetched before waiting for the accumulator
        const bool refresh = vec && (!BRES || mt == 0);
        float s_pre = p.alpha, t_pre = 0.f;
        if (refresh && tid_e < BN) {
          const long col = (long)n_tile * BN + tid_e;
          float g = g_u, bv = 0.f;
          if (col < p.N) {
            if (gcol) g = p.gate_tanh ? tanhf(p.gate[col]) : p.gate[col];
            if (bcol) bv = p.bias[z1 * p.bias_b1 + z2 * p.bias_b2 + col];
          }
          s_pre = p.alpha * g;
          t_pre = bscale * bv * g;
        }
        const long grow_idx = (long)(m0 + mt) * BM + row;
        float g_r = g_u, b_r = 0.f;
        if (grow_idx < p.M) {
          if (grow) g_r = p.gate_tanh ? tanhf(p.gate[grow_idx]) : p.gate[grow_idx];
          if (brow) b_r = p.bias[z1 * p.bias_b1 + z2 * p.bias_b2 + grow_idx];
        }
        const float s_u = p.alpha * g_r, t_u = bscale * b_r * g_r;
        mbar_wait(&tmem_full_bar[buf], (lt >> 1) & 1);
        tc_fence_after();
        if (refresh) {  // every epilogue warp left the previous tile's window loop at its closing barrier
          if (tid_e < BN) {
            s_vec[tid_e] = s_pre;
            t_vec[tid_e] = t_pre;
          }
          asm volatile("bar.sync 1, 512;" ::: "memory");
        }
        const long nl = p.N - (long)n_tile * BN;
        const int nwin = nl >= BN ? 4 : (int)((nl + 63) >> 6);
        const uint32_t t_acc = tmem_base + (uint32_t)(buf * BN) + lane_addr + (uint32_t)(part * 16);
        uint32_t ra[16], rb[16];
        tmem_ld_32x16(t_acc, ra);  // window 0
        for (int w = 0; w < nwin; ++w, ++wcount) {
          uint32_t (&r)[16] = (w & 1) ? rb : ra;
          uint32_t (&rn)[16] = (w & 1) ? ra : rb;
          tmem_ld_wait_dep(r);
          if (w + 1 < nwin) tmem_ld_32x16(t_acc + (uint32_t)((w + 1) * 64), rn);  // next window in flight
          const int c0 = w * 64 + part * 16;  // this thread's 16 columns inside the tile
          float v[16];
          if (vec) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 sv = lds128f(s_addr + (uint32_t)(c0 * 4 + k * 16));
              const float4 tv = lds128f(t_addr + (uint32_t)(c0 * 4 + k * 16));
              ffma2v(v[4 * k], v[4 * k + 1], __uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), sv.x, sv.y, tv.x, tv.y);
              ffma2v(v[4 * k + 2], v[4 * k + 3], __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]), sv.z, sv.w, tv.z,
                     tv.w);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              ffma2(v[2 * k], v[2 * k + 1], __uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]), s_u, t_u, t_u);
          }
          __half2 h[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
          if (clampv > 0.f) {
            const __half2 hi = __float2half2_rn(clampv), lo = __float2half2_rn(-clampv);
#pragma unroll
            for (int k = 0; k < 8; ++k) h[k] = __hmax2(__hmin2(h[k], hi), lo);
          }
          const uint32_t blk = smem_u32(stg + (wcount & 1) * (BM * 128)) + row * 128;
          const uint32_t* hv = reinterpret_cast<const uint32_t*>(h);
          const int j0 = part * 2;
          sts128(blk + (((j0) ^ sw) << 4), hv[0], hv[1], hv[2], hv[3]);
          sts128(blk + (((j0 + 1) ^ sw) << 4), hv[4], hv[5], hv[6], hv[7]);
          if (w == nwin - 1) {  // the accumulator is in registers / staged: release it to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
          }
          fence_proxy_async();
          // the store issued one window ago used the OTHER staging tile, which the next window rewrites after this barrier
          if (issuer) tma_store_wait_read_all();
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (issuer) {
            const int cz1 = p.nb1 == 1 ? 0 : z1, cz2 = p.nb2 == 1 ? 0 : z2;
            tma_store_4d(&tma_c, stg + (wcount & 1) * (BM * 128), (int)((long)n_tile * BN + w * 64), (m0 + mt) * BM, cz1, cz2);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_read_all();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// Host side: same operand / output maps and span distribution as launch_tcp<256, STAGES, BRES>.
template <int STAGES, bool BRES>
static int launch_plain16(const GemmP& p0, cudaStream_t st) {
  constexpr int BN = 256;
  using Cfg = TcpCfg<BN, STAGES, BRES>;
  GemmP p = p0;
  CUtensorMap ma, mb, mcm;
  int rc = make_operand_map(&ma, p.A, p.M, p.K, p.lda, p.nb1, p.a_b1, p.nb2, p.a_b2, BM, &p.a_bcast1, &p.a_bcast2);
  if (rc) return rc;
  rc = make_operand_map(&mb, p.B, p.N, p.K, p.ldb, p.nb1, p.b_b1, p.nb2, p.b_b2, BN, &p.b_bcast1, &p.b_bcast2);
  if (rc) return rc;
  if (!(can_tma_store(p, BN, true) && fast_epilogue_ok(p) && p.c_dtype == MQDET_F16 && p.act == MQDET_ACT_NONE && !p.R)) {
    set_error("gemm_plain16: not a plain fp16 TMA-store product");
    return MQDET_ERR_ARG;
  }
  p.use_tma_store = 1;
  p.fast_epi = 1;
  rc = make_output_map(&mcm, p);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(gemm_plain16_kernel<STAGES, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    return MQDET_ERR_CUDA;
  }
  const int tm = cdiv(p.M, BM), tn = cdiv(p.N, BN);
  const long Z = (long)p.nb1 * p.nb2;
  const long total = (long)tm * tn * Z;
  const long sms = num_sms();
  const long g = total < sms ? total : sms;
  const long share = (total + g - 1) / g;
  const int mc = !BRES ? 1 : (tn == 1 ? tm : (int)(share < 1 ? 1 : (share > tm ? tm : share)));
  gemm_plain16_kernel<STAGES, BRES><<<(int)g, 640, Cfg::SMEM_BYTES, st>>>(ma, mb, mcm, p, tm, tn, (int)total, mc);
  return check_launch("gemm_plain16_kernel");
}

}  // namespace mqdet

// Entry point for round-2 experiments (same argument struct as mqdet_gemm_f16); K <= 256 -> B-resident variant.
extern "C" int mqdet_exp_gemm_plain16(const mqdet_gemm_args* a, void* stream) {
  mqdet::GemmP p;
  memset(&p, 0, sizeof(p));
  p.A = (const __half*)a->A;
  p.B = (const __half*)a->B;
  p.M = a->M; p.N = a->N; p.K = a->K; p.lda = a->lda; p.ldb = a->ldb;
  p.nb1 = (int)a->nb1; p.nb2 = (int)a->nb2;
  p.a_b1 = a->a_b1; p.a_b2 = a->a_b2; p.b_b1 = a->b_b1; p.b_b2 = a->b_b2;
  p.C = a->C; p.c_dtype = a->c_dtype; p.ldc = a->ldc; p.c_b1 = a->c_b1; p.c_b2 = a->c_b2;
  p.alpha = a->alpha; p.scale_after_bias = a->scale_after_bias;
  p.bias = a->bias; p.bias_mode = a->bias ? a->bias_mode : MQDET_VEC_NONE;
  p.bias_b1 = a->bias_b1; p.bias_b2 = a->bias_b2;
  p.act = a->act; p.clamp = a->clamp;
  p.gate = a->gate; p.gate_mode = a->gate ? a->gate_mode : MQDET_VEC_NONE; p.gate_tanh = a->gate_tanh;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.K <= mqdet::BRES_KB * mqdet::BK) return mqdet::launch_plain16<3, true>(p, st);
  return mqdet::launch_plain16<3, false>(p, st);
}
