// mqdet_b200 — multi-scale deformable attention forward (GroundingDINO encoder / decoder cross-attention).
//
// Reference: groundingdino_new/models/GroundingDINO/ms_deform_attn.py — MultiScaleDeformableAttention.forward :236-352 and its
// pure-torch core multi_scale_deformable_attn_pytorch :93-133 (grid_sample bilinear, zeros padding, align_corners=False); the
// reference's own CUDA kernel is csrc_groundingdino/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-298 (a scalar grid-stride loop
// that does not compile against torch 2.11).
//
// Fused here per (image, query, head) — one WARP, lane == channel of the 32-wide head:
//   softmax over the L*P attention logits (lanes 0..L*P-1, shuffle reductions)         (:290-299)
//   sampling location = reference point + offset / (W_l, H_l)  (2-d reference points)   (:303-308)
//                     = reference centre + offset / P * reference size * 0.5 (4-d boxes) (:309-316)
//   bilinear sample of value[b, level rows, h*32 + lane] at (x*W - 0.5, y*H - 0.5), zero outside, weighted sum
// HBM/L2-bound gather: every corner read is one coalesced 64-byte row segment of the fp16 value tensor.
#include "common.cuh"
#include "../../include/mqdet_b200.h"

namespace mqdet {

struct MsdaLevels {
  int n;
  int H[MQDET_MAX_LEVELS], W[MQDET_MAX_LEVELS], off[MQDET_MAX_LEVELS];
};

// value [B][Nv][heads*32] f16; proj [B*Q][proj_ld] f32 holding the sampling offsets at column h*L*P*2 + (l*P + p)*2 + {0,1}
// and the attention logits at column aw_col0 + h*L*P + l*P + p; ref [B][Q][L][ref_dim] f32; out [B*Q][heads*32].
template <typename OutT>
__global__ void __launch_bounds__(256) ms_deform_attn_kernel(const __half* __restrict__ value, const float* __restrict__ proj,
                                                             int proj_ld, int aw_col0, const float* __restrict__ ref, int ref_dim,
                                                             MsdaLevels lv, int B, int Q, int Nv, int heads, int P,
                                                             OutT* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long unit = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= (long)B * Q * heads) return;
  const int h = (int)(unit % heads);
  const long bq = unit / heads;
  const int b = (int)(bq / Q);
  const int L = lv.n, LP = L * P;
  const float* pr = proj + bq * proj_ld;
  // softmax over the L*P logits of this (query, head)
  float logit = lane < LP ? pr[aw_col0 + h * LP + lane] : -INFINITY;
  const float mx = warp_max(logit);
  float e = lane < LP ? __expf(logit - mx) : 0.f;
  const float aw = e / warp_sum(e);
  // this lane's own sampling point (lane < LP): offset (x, y)
  float ox = 0.f, oy = 0.f;
  if (lane < LP) {
    ox = pr[(h * LP + lane) * 2];
    oy = pr[(h * LP + lane) * 2 + 1];
  }
  const __half* vb = value + (long)b * Nv * heads * 32 + h * 32 + lane;
  const int C = heads * 32;
  float acc = 0.f;
  for (int i = 0; i < LP; ++i) {
    const int l = i / P;
    const float w_i = __shfl_sync(0xffffffffu, aw, i);
    const float ox_i = __shfl_sync(0xffffffffu, ox, i), oy_i = __shfl_sync(0xffffffffu, oy, i);
    const float* rp = ref + (bq * L + l) * ref_dim;
    const int Hl = lv.H[l], Wl = lv.W[l];
    float lx, ly;
    if (ref_dim == 2) {
      lx = rp[0] + ox_i / (float)Wl;
      ly = rp[1] + oy_i / (float)Hl;
    } else {
      lx = rp[0] + ox_i / (float)P * rp[2] * 0.5f;
      ly = rp[1] + oy_i / (float)P * rp[3] * 0.5f;
    }
    // grid_sample(align_corners=False): pixel = loc * size - 0.5
    const float x = lx * (float)Wl - 0.5f, y = ly * (float)Hl - 0.5f;
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float ax = x - xf, ay = y - yf;
    const __half* vl = vb + (long)lv.off[l] * C;
    float s = 0.f;
    if (y0 >= 0 && y0 < Hl) {
      if (x0 >= 0 && x0 < Wl) s += (1.f - ay) * (1.f - ax) * __half2float(vl[((long)y0 * Wl + x0) * C]);
      if (x0 + 1 >= 0 && x0 + 1 < Wl) s += (1.f - ay) * ax * __half2float(vl[((long)y0 * Wl + x0 + 1) * C]);
    }
    if (y0 + 1 >= 0 && y0 + 1 < Hl) {
      if (x0 >= 0 && x0 < Wl) s += ay * (1.f - ax) * __half2float(vl[((long)(y0 + 1) * Wl + x0) * C]);
      if (x0 + 1 >= 0 && x0 + 1 < Wl) s += ay * ax * __half2float(vl[((long)(y0 + 1) * Wl + x0 + 1) * C]);
    }
    acc = fmaf(w_i, s, acc);
  }
  if (sizeof(OutT) == 2)
    reinterpret_cast<__half*>(out)[bq * C + h * 32 + lane] = __float2half_rn(acc);
  else
    reinterpret_cast<float*>(out)[bq * C + h * 32 + lane] = acc;
}

// Second mapping (used when L*P is a multiple of 8): the warp of a (query, head) is split into 8 sample groups x 4 channel quads —
// lane = g*4 + c, g = sample group, c = channels c*8 .. c*8+7.  Every lane owns ONE sampling point per pass (its own offset, weight
// and bilinear coefficients) and reads each of its four corners as one 16-byte vector, so a pass issues 4 load instructions that
// together cover 8 points x 4 corners x 64 bytes; L*P = 16 points take 2 passes = 8 load instructions per lane instead of the 64
// two-byte loads of the lane-per-channel mapping above.  The 8 partial sums of a channel are combined by three xor-shuffles.
template <typename OutT>
__global__ void __launch_bounds__(256) ms_deform_attn_v2_kernel(const __half* __restrict__ value, const float* __restrict__ proj,
                                                                int proj_ld, int aw_col0, const float* __restrict__ ref, int ref_dim,
                                                                MsdaLevels lv, int B, int Q, int Nv, int heads, int P,
                                                                OutT* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long unit = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= (long)B * Q * heads) return;
  const int h = (int)(unit % heads);
  const long bq = unit / heads;
  const int b = (int)(bq / Q);
  const int L = lv.n, LP = L * P;
  const float* pr = proj + bq * proj_ld;
  const float logit = lane < LP ? pr[aw_col0 + h * LP + lane] : -INFINITY;
  const float mx = warp_max(logit);
  const float e = lane < LP ? __expf(logit - mx) : 0.f;
  const float aw = e / warp_sum(e);
  const int g = lane >> 2, c = lane & 3;
  const int C = heads * 32;
  const __half* vb = value + (long)b * Nv * C + h * 32 + c * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int i0 = 0; i0 < LP; i0 += 8) {
    const int i = i0 + g;                       // this lane's sampling point
    const int l = i / P;
    const float w_i = __shfl_sync(0xffffffffu, aw, i);
    const float ox = pr[(h * LP + i) * 2], oy = pr[(h * LP + i) * 2 + 1];
    const float* rp = ref + (bq * L + l) * ref_dim;
    const int Hl = lv.H[l], Wl = lv.W[l];
    float lx, ly;
    if (ref_dim == 2) {
      lx = rp[0] + ox / (float)Wl;
      ly = rp[1] + oy / (float)Hl;
    } else {
      lx = rp[0] + ox / (float)P * rp[2] * 0.5f;
      ly = rp[1] + oy / (float)P * rp[3] * 0.5f;
    }
    const float x = lx * (float)Wl - 0.5f, y = ly * (float)Hl - 0.5f;
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float ax = x - xf, ay = y - yf;
    const __half* vl = vb + (long)lv.off[l] * C;
    const bool okx0 = x0 >= 0 && x0 < Wl, okx1 = x0 + 1 >= 0 && x0 + 1 < Wl;
    const bool oky0 = y0 >= 0 && y0 < Hl, oky1 = y0 + 1 >= 0 && y0 + 1 < Hl;
    const float wt[4] = {oky0 && okx0 ? w_i * (1.f - ay) * (1.f - ax) : 0.f, oky0 && okx1 ? w_i * (1.f - ay) * ax : 0.f,
                         oky1 && okx0 ? w_i * ay * (1.f - ax) : 0.f, oky1 && okx1 ? w_i * ay * ax : 0.f};
    // clamped addresses: every load is issued unconditionally (a zero weight discards what an out-of-range corner reads)
    const int xc0 = min(max(x0, 0), Wl - 1), xc1 = min(max(x0 + 1, 0), Wl - 1);
    const int yc0 = min(max(y0, 0), Hl - 1), yc1 = min(max(y0 + 1, 0), Hl - 1);
    uint4 u[4];
    u[0] = *reinterpret_cast<const uint4*>(vl + ((long)yc0 * Wl + xc0) * C);
    u[1] = *reinterpret_cast<const uint4*>(vl + ((long)yc0 * Wl + xc1) * C);
    u[2] = *reinterpret_cast<const uint4*>(vl + ((long)yc1 * Wl + xc0) * C);
    u[3] = *reinterpret_cast<const uint4*>(vl + ((long)yc1 * Wl + xc1) * C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2* hv = reinterpret_cast<const __half2*>(&u[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(hv[k]);
        acc[2 * k] = fmaf(wt[j], f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(wt[j], f.y, acc[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 4);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
  }
  if (g == 0) {
    const long o = bq * C + h * 32 + c * 8;
    if (sizeof(OutT) == 2) {
      __half2 hv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) hv[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + o) = *reinterpret_cast<uint4*>(hv);
    } else {
      float* of = reinterpret_cast<float*>(out) + o;
      *reinterpret_cast<float4*>(of) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(of + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

}  // namespace mqdet

using namespace mqdet;

extern "C" int mqdet_ms_deform_attn(const void* value, const float* proj, int64_t proj_ld, int64_t aw_col0, const float* ref,
                                    int64_t ref_dim, const int32_t* level_hw, int64_t nlev, int64_t B, int64_t Q, int64_t heads,
                                    int64_t head_dim, int64_t points, void* out, int out_dtype, void* stream) {
  MQ_REQUIRE(value && proj && ref && level_hw && out, "ms_deform_attn: null pointer");
  MQ_REQUIRE(head_dim == 32, "ms_deform_attn: head dim must be 32 (embed 256 / 8 heads), got %ld", (long)head_dim);
  MQ_REQUIRE(nlev >= 1 && nlev <= MQDET_MAX_LEVELS && points >= 1 && nlev * points <= 32, "ms_deform_attn: levels * points must be <= 32");
  MQ_REQUIRE(ref_dim == 2 || ref_dim == 4, "ms_deform_attn: reference points must have 2 or 4 coordinates");
  MQ_REQUIRE(B >= 1 && Q >= 1 && heads >= 1, "ms_deform_attn: empty problem");
  MQ_REQUIRE(out_dtype == MQDET_F16 || out_dtype == MQDET_F32, "ms_deform_attn: bad out_dtype");
  MsdaLevels lv;
  lv.n = (int)nlev;
  int off = 0;
  for (int l = 0; l < nlev; ++l) {
    lv.H[l] = level_hw[2 * l];
    lv.W[l] = level_hw[2 * l + 1];
    lv.off[l] = off;
    MQ_REQUIRE(lv.H[l] > 0 && lv.W[l] > 0, "ms_deform_attn: bad level table");
    off += lv.H[l] * lv.W[l];
  }
  const long units = B * Q * heads;
  const unsigned grid = (unsigned)((units + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v2 = ((nlev * points) % 8) == 0 && (((uintptr_t)value) & 15) == 0 && (((uintptr_t)out) & 15) == 0;
#define MQ_MSDA_ARGS (const __half*)value, proj, (int)proj_ld, (int)aw_col0, ref, (int)ref_dim, lv, (int)B, (int)Q, off, (int)heads, (int)points
  if (out_dtype == MQDET_F16) {
    if (v2) ms_deform_attn_v2_kernel<__half><<<grid, 256, 0, st>>>(MQ_MSDA_ARGS, (__half*)out);
    else ms_deform_attn_kernel<__half><<<grid, 256, 0, st>>>(MQ_MSDA_ARGS, (__half*)out);
  } else {
    if (v2) ms_deform_attn_v2_kernel<float><<<grid, 256, 0, st>>>(MQ_MSDA_ARGS, (float*)out);
    else ms_deform_attn_kernel<float><<<grid, 256, 0, st>>>(MQ_MSDA_ARGS, (float*)out);
  }
#undef MQ_MSDA_ARGS
  return check_launch("ms_deform_attn_kernel");
}
