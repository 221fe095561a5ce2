// mqdet_b200 — Gated Class-scalable Perceiver (GCP) kernels.
//
// Reference: maskrcnn_benchmark/modeling/language_backbone/modeling_bert_new.py
//   get_index_with_padding_batch :40-63, _construct_sparse_inputs :162-184,
//   MaskedCrossAttention.forward (sparse) :186-248, GatedCrossAttentionBlock.forward :347-374.
//
// The reference gathers the <=S vision queries of every text token and re-projects them to K/V per
// token (B*T*S rows).  K/V only depend on the unique query, so here they are projected ONCE for the
// V+1 rows (row V = the zero padding slot pushed through norm_kv) and the attention kernel gathers the
// projected rows: one warp per token, 8 heads x 64 dims = 16 contiguous dims per lane, 4 lanes per head,
// warp-shuffle reductions for q.k, softmax over S in registers.
#include "common.cuh"
#include "../../include/mqdet_b200.h"

namespace mqdet {

constexpr int GCP_MAX_S = 16;

__device__ __forceinline__ void ld16h(const __half* p, float (&f)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 x = __half22float2(ha[i]);
    float2 y = __half22float2(hb[i]);
    f[2 * i] = x.x; f[2 * i + 1] = x.y;
    f[8 + 2 * i] = y.x; f[8 + 2 * i + 1] = y.y;
  }
}

// Generic in heads/dim as long as H*Dh == 32 * 16 (8x64 GCP) -- lanes_per_head = Dh/16.
__global__ void __launch_bounds__(256) gcp_sparse_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ kv,
                                                              const int* __restrict__ idx, __half* __restrict__ out,
                                                              long BT, int T, int V, int S, int lanes_per_head) {
  const long tok = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= BT) return;
  const int b = (int)(tok / T);
  const int inner = 32 * 16;  // H*Dh
  const __half* kvb = kv + (long)b * (V + 1) * (2 * inner);
  float qf[16];
  ld16h(q + tok * inner + lane * 16, qf);
  float sim[GCP_MAX_S];
  int id[GCP_MAX_S];
#pragma unroll
  for (int s = 0; s < GCP_MAX_S; ++s) {
    if (s < S) {
      id[s] = idx[tok * S + s];
      float kf[16];
      ld16h(kvb + (long)id[s] * (2 * inner) + lane * 16, kf);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) d = fmaf(qf[i], kf[i], d);
      for (int o = 1; o < lanes_per_head; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      // additive -1e4 on padding slots (":219-223", "for half")
      sim[s] = d + (id[s] == V ? -1e4f : 0.f);
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < GCP_MAX_S; ++s)
    if (s < S) mx = fmaxf(mx, sim[s]);
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < GCP_MAX_S; ++s)
    if (s < S) {
      sim[s] = expf(sim[s] - mx);
      den += sim[s];
    }
  const float inv = 1.f / den;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int s = 0; s < GCP_MAX_S; ++s) {
    if (s < S) {
      // probabilities of padding slots are zeroed AFTER the softmax, without renormalisation (:227-231)
      const float p = (id[s] == V) ? 0.f : sim[s] * inv;
      float vf[16];
      ld16h(kvb + (long)id[s] * (2 * inner) + inner + lane * 16, vf);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
    }
  }
  __half2 h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
  uint4* dst = reinterpret_cast<uint4*>(out + tok * inner + lane * 16);
  dst[0] = *reinterpret_cast<uint4*>(&h[0]);
  dst[1] = *reinterpret_cast<uint4*>(&h[4]);
}

// g = tanh(h1 . w2); x1 = s*g + x; ln = LN(x1).  One warp per row.
__global__ void __launch_bounds__(256) gcp_gate_residual_ln_kernel(const __half* __restrict__ h1, const float* __restrict__ w2,
                                                                   int Dg, const float* __restrict__ s,
                                                                   const float* __restrict__ x,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float eps, long rows,
                                                                   int D, float* __restrict__ x1_out,
                                                                   __half* __restrict__ ln_out, float* __restrict__ gate_out) {
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float d = 0.f;
  for (int i = lane; i < Dg; i += 32) d = fmaf(__half2float(h1[row * Dg + i]), w2[i], d);
  const float g = tanhf(warp_sum(d));
  if (gate_out && lane == 0) gate_out[row] = g;
  const float* sr = s + row * D;
  const float* xr = x + row * D;
  float sum = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = fmaf(sr[i], g, xr[i]);
    x1_out[row * D + i] = v;
    sum += v;
  }
  const float mean = warp_sum(sum) / D;
  float var = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = fmaf(sr[i], g, xr[i]) - mean;
    var += v * v;
  }
  const float rstd = rsqrtf(warp_sum(var) / D + eps);
  for (int i = lane; i < D; i += 32) {
    const float v = fmaf(sr[i], g, xr[i]);
    ln_out[row * D + i] = __float2half_rn((v - mean) * rstd * gamma[i] + beta[i]);
  }
}

// idx[b,t,0:S] = ascending v with mask[b,v,t] != 0, padded with V.
__global__ void gcp_build_index_kernel(const float* __restrict__ mask, int B, int V, int T, int S, int* __restrict__ idx,
                                       int* __restrict__ counts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * T) return;
  const int b = (int)(i / T), t = (int)(i % T);
  const float* m = mask + (long)b * V * T + t;
  int n = 0, total = 0;
  for (int v = 0; v < V; ++v) {
    if (m[(long)v * T] != 0.f) {
      if (n < S) idx[i * S + n++] = v;
      ++total;
    }
  }
  for (; n < S; ++n) idx[i * S + n] = V;
  if (counts) counts[i] = total;
}

}  // namespace mqdet

using namespace mqdet;

extern "C" int mqdet_gcp_sparse_attn(const void* q, const void* kv, const int32_t* idx, void* out, int64_t B, int64_t T,
                                     int64_t V, int64_t S, int64_t H, int64_t Dh, void* stream) {
  MQ_REQUIRE(q && kv && idx && out, "gcp_sparse_attn: null pointer");
  MQ_REQUIRE(H * Dh == 512 && (Dh % 16) == 0 && Dh <= 512, "gcp_sparse_attn: H*Dh must be 512 (got %ld x %ld)", (long)H,
             (long)Dh);
  MQ_REQUIRE(S >= 1 && S <= GCP_MAX_S, "gcp_sparse_attn: S=%ld out of range [1,%d]", (long)S, GCP_MAX_S);
  const int lph = (int)(Dh / 16);
  MQ_REQUIRE((lph & (lph - 1)) == 0, "gcp_sparse_attn: Dh/16 must be a power of two");
  const long BT = B * T;
  const int wpb = 8;
  gcp_sparse_attn_kernel<<<cdiv(BT, wpb), wpb * 32, 0, (cudaStream_t)stream>>>((const __half*)q, (const __half*)kv, idx,
                                                                             (__half*)out, BT, (int)T, (int)V, (int)S, lph);
  return check_launch("gcp_sparse_attn_kernel");
}

extern "C" int mqdet_gcp_gate_residual_ln(const void* h1, const float* w2, int64_t Dg, const float* s, const float* x,
                                          const float* gamma, const float* beta, float eps, int64_t rows, int64_t D,
                                          float* x1_out, void* ln_out16, float* gate_out, void* stream) {
  MQ_REQUIRE(h1 && w2 && s && x && gamma && beta && x1_out && ln_out16, "gcp_gate_residual_ln: null pointer");
  const int wpb = 8;
  gcp_gate_residual_ln_kernel<<<cdiv(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      (const __half*)h1, w2, (int)Dg, s, x, gamma, beta, eps, rows, (int)D, x1_out, (__half*)ln_out16, gate_out);
  return check_launch("gcp_gate_residual_ln_kernel");
}

extern "C" int mqdet_gcp_build_index(const float* mask, int64_t B, int64_t V, int64_t T, int64_t S, int32_t* idx,
                                     int32_t* counts_out, void* stream) {
  MQ_REQUIRE(mask && idx && B > 0 && V > 0 && T > 0 && S > 0, "gcp_build_index: bad args");
  const long n = B * T;
  gcp_build_index_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(mask, (int)B, (int)V, (int)T, (int)S, idx,
                                                                        counts_out);
  return check_launch("gcp_build_index_kernel");
}
