// mqdet_b200 — multi-label NMS entirely on the device (no D2H of the suppression mask, no host scan).
//
// Reference: maskrcnn_benchmark/csrc/cuda/ml_nms.cu (devIoU :15-26, ml_nms_kernel :28-76, host scan :129-140,
// ascending original indices :145-149) and the top-k cut of rpn/inference.py:757-767.
// Same arithmetic (fp32, +1 pixel widths, label-gated, strict '>' threshold) so kept indices are bit-identical.
//
//   1. argsort_desc_kernel  : single-CTA bitonic sort of (score desc, index asc) -> order   [n <= 16384]
//   2. nms_mask_kernel      : 64x64 blocks of the upper-triangular suppression bit matrix (sorted order)
//   3. nms_scan_kernel      : single CTA; per 64-box block, one thread resolves the diagonal word serially,
//                             all threads OR the kept rows into the removed-set words of later blocks
//   4. nms_compact_kernel   : flags in ORIGINAL index space -> ascending kept indices (+ optional top-k cut)
#include "common.cuh"
#include "../../include/mqdet_b200.h"

namespace mqdet {

constexpr int NMS_MAX_N = 16384;
constexpr int NMS_MAX_BLOCKS = NMS_MAX_N / 64;

__device__ __forceinline__ float dev_iou(const float* a, const float* b) {
  if (a[5] != b[5]) return 0.0f;
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

// keys: score descending, ties by ascending original index (== a stable descending sort).
__global__ void __launch_bounds__(1024) argsort_desc_kernel(const float* __restrict__ scores, int n, long long* __restrict__ order,
                                                            const int* __restrict__ n_arr, int n_max) {
  extern __shared__ unsigned long long keys[];
  if (n_arr) {  // batched: one CTA per image, rows of n_max
    n = min(n_arr[blockIdx.x], n_max);
    scores += (long)blockIdx.x * n_max;
    order += (long)blockIdx.x * n_max;
    if (n == 0) return;
  }
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    unsigned long long k = ~0ull;  // sentinels sort last
    if (i < n) {
      unsigned int u = __float_as_uint(scores[i]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-sortable
      u = ~u;                                          // -> descending
      k = ((unsigned long long)u << 32) | (unsigned int)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = ((i & k) == 0);
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = (long long)(keys[i] & 0xffffffffu);
}

// gathers boxes into sorted [n,6] rows (x1,y1,x2,y2,score,label) — the layout ml_nms.cu works on.
__global__ void nms_gather_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                  const float* __restrict__ labels, const long long* __restrict__ order, int n,
                                  float* __restrict__ sorted, const int* __restrict__ n_arr, int n_max) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_arr) {
    const int b = blockIdx.y;
    n = min(n_arr[b], n_max);
    boxes += (long)b * n_max * 4; scores += (long)b * n_max; labels += (long)b * n_max; order += (long)b * n_max;
    sorted += (long)b * n_max * 6;
  }
  if (i >= n) return;
  long long o = order[i];
  sorted[i * 6 + 0] = boxes[o * 4 + 0];
  sorted[i * 6 + 1] = boxes[o * 4 + 1];
  sorted[i * 6 + 2] = boxes[o * 4 + 2];
  sorted[i * 6 + 3] = boxes[o * 4 + 3];
  sorted[i * 6 + 4] = scores[o];
  sorted[i * 6 + 5] = labels[o];
}

__global__ void __launch_bounds__(64) nms_mask_kernel(int n, float thresh, const float* __restrict__ sorted,
                                                      unsigned long long* __restrict__ mask, const int* __restrict__ n_arr,
                                                      int n_max) {
  const int row_start = blockIdx.y, col_start = blockIdx.x;
  if (n_arr) {
    const int b = blockIdx.z;
    n = min(n_arr[b], n_max);
    sorted += (long)b * n_max * 6;
    mask += (long)b * n_max * ((n_max + 63) / 64);
  }
  const int col_blocks = (n + 63) / 64;
  if (col_start < row_start || row_start >= col_blocks || col_start >= col_blocks) {  // lower triangle is never read
    return;
  }
  const int row_size = min(n - row_start * 64, 64);
  const int col_size = min(n - col_start * 64, 64);
  __shared__ float bb[64 * 6];
  if ((int)threadIdx.x < col_size) {
#pragma unroll
    for (int k = 0; k < 6; ++k) bb[threadIdx.x * 6 + k] = sorted[(64 * col_start + threadIdx.x) * 6 + k];
  }
  __syncthreads();
  if ((int)threadIdx.x < row_size) {
    const int cur = 64 * row_start + threadIdx.x;
    const float* cb = sorted + cur * 6;
    unsigned long long t = 0;
    int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
    for (int i = start; i < col_size; ++i)
      if (dev_iou(cb, bb + i * 6) > thresh) t |= 1ULL << i;
    mask[(long)cur * col_blocks + col_start] = t;
  }
}

// Greedy scan in sorted order. flags[orig] = 1 for kept boxes; kth_score = score of the max_det-th kept box.
__global__ void __launch_bounds__(256) nms_scan_kernel(int n, const unsigned long long* __restrict__ mask,
                                                       const long long* __restrict__ order, const float* __restrict__ sorted,
                                                       int max_det, unsigned char* __restrict__ flags,
                                                       float* __restrict__ kth_score, int* __restrict__ n_kept_sorted,
                                                       const int* __restrict__ n_arr, int n_max) {
  if (n_arr) {
    const int b = blockIdx.x;
    n = min(n_arr[b], n_max);
    mask += (long)b * n_max * ((n_max + 63) / 64);
    order += (long)b * n_max; sorted += (long)b * n_max * 6; flags += (long)b * n_max;
    kth_score += b; n_kept_sorted += b;
  }
  __shared__ unsigned long long remv[NMS_MAX_BLOCKS];
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long keepbits;
  __shared__ int kept_so_far;
  const int col_blocks = (n + 63) / 64;
  for (int i = threadIdx.x; i < col_blocks; i += blockDim.x) remv[i] = 0;
  if (threadIdx.x == 0) {
    kept_so_far = 0;
    *kth_score = -INFINITY;
  }
  __syncthreads();
  __shared__ int kept_list[64];
  __shared__ int kept_cnt;
  for (int blk = 0; blk < col_blocks; ++blk) {
    const int base = blk * 64;
    const int size = min(n - base, 64);
    // Early exit (top-k cut only): once max_det boxes are kept, a later box can pass the `score >= kth kept score` cut of
    // rpn/inference.py:759-767 only with a score equal to it; scores are sorted descending, so when the first score of a
    // block is already below the threshold nothing after it can appear in the output (kept or not).
    if (max_det > 0 && kept_so_far >= max_det && sorted[base * 6 + 4] < *kth_score) break;
    if ((int)threadIdx.x < size) diag[threadIdx.x] = mask[(long)(base + threadIdx.x) * col_blocks + blk];
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long r = remv[blk], kb = 0;
      int cnt = 0;
      for (int t = 0; t < size; ++t) {
        if (!(r & (1ULL << t))) {
          kb |= 1ULL << t;
          r |= diag[t];
          kept_list[cnt++] = t;
          ++kept_so_far;
          if (max_det > 0 && kept_so_far == max_det) *kth_score = sorted[(base + t) * 6 + 4];
        }
      }
      keepbits = kb;
      kept_cnt = cnt;
    }
    __syncthreads();
    const unsigned long long kb = keepbits;
    const int cnt = kept_cnt;
    if ((int)threadIdx.x < size && (kb & (1ULL << threadIdx.x))) flags[order[base + threadIdx.x]] = 1;
    // OR the kept rows into the removed words of later column blocks: 8 independent loads in flight per thread
    for (int j = blk + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
      unsigned long long r = remv[j];
      const unsigned long long* mrow = mask + (long)base * col_blocks + j;
      int k = 0;
      for (; k + 8 <= cnt; k += 8) {
        unsigned long long w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = mrow[(long)kept_list[k + u] * col_blocks];
#pragma unroll
        for (int u = 0; u < 8; ++u) r |= w[u];
      }
      for (; k < cnt; ++k) r |= mrow[(long)kept_list[k] * col_blocks];
      remv[j] = r;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_kept_sorted = kept_so_far;
}

// Ascending compaction of flagged original indices; if more than max_det were kept, apply the
// "score >= kth largest kept score" cut of rpn/inference.py:759-767 (ties keep extras).
__global__ void __launch_bounds__(1024) nms_compact_kernel(int n, const unsigned char* __restrict__ flags,
                                                           const float* __restrict__ scores, int max_det,
                                                           const float* __restrict__ kth_score,
                                                           const int* __restrict__ n_kept_sorted,
                                                           long long* __restrict__ keep_out, int* __restrict__ num_keep,
                                                           const int* __restrict__ n_arr, int n_max) {
  if (n_arr) {
    const int b = blockIdx.x;
    n = min(n_arr[b], n_max);
    flags += (long)b * n_max; scores += (long)b * n_max; keep_out += (long)b * n_max;
    kth_score += b; n_kept_sorted += b; num_keep += b;
  }
  __shared__ int warp_tot[32];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const bool cut = max_det > 0 && *n_kept_sorted > max_det;
  const float thr = *kth_score;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int f = 0;
    if (i < n && flags[i]) f = cut ? (scores[i] >= thr) : 1;
    const unsigned ball = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int prefix = __popc(ball & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(ball);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int start = running;
    if (f) keep_out[start + woff + prefix] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
      running = start + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = running;
}

}  // namespace mqdet

using namespace mqdet;

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int mqdet_argsort_desc(const float* scores, int64_t n, int64_t* order, void* stream) {
  MQ_REQUIRE(scores && order, "argsort_desc: null pointer");
  MQ_REQUIRE(n >= 0 && n <= NMS_MAX_N, "argsort_desc: n=%ld exceeds %d", (long)n, NMS_MAX_N);
  if (n == 0) return MQDET_OK;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&argsort_desc_kernel), NMS_MAX_N * 8)) return rc;
  argsort_desc_kernel<<<1, 1024, (size_t)np2 * 8, (cudaStream_t)stream>>>(scores, (int)n, (long long*)order, nullptr, 0);
  return check_launch("argsort_desc_kernel");
}

extern "C" int64_t mqdet_ml_nms_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  const size_t cb = (size_t)(n + 63) / 64;
  return (int64_t)(align256((size_t)n * 6 * 4) + align256((size_t)n * cb * 8) + align256((size_t)n) + 256);
}

extern "C" int mqdet_ml_nms(const float* boxes, const float* scores, const float* labels, const int64_t* order, int64_t n,
                            float thresh, int64_t max_det, int64_t* keep_out, int32_t* num_keep, void* workspace,
                            void* stream) {
  MQ_REQUIRE(num_keep, "ml_nms: null num_keep");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {  // reference returns an empty tensor (csrc/ml_nms.h:19-20)
    cudaMemsetAsync(num_keep, 0, sizeof(int32_t), st);
    return MQDET_OK;
  }
  MQ_REQUIRE(boxes && scores && labels && order && keep_out && workspace, "ml_nms: null pointer");
  MQ_REQUIRE(n <= NMS_MAX_N, "ml_nms: n=%ld exceeds %d", (long)n, NMS_MAX_N);
  const int cb = (int)((n + 63) / 64);
  uint8_t* ws = (uint8_t*)workspace;
  float* sorted = (float*)ws;
  ws += align256((size_t)n * 6 * 4);
  unsigned long long* mask = (unsigned long long*)ws;
  ws += align256((size_t)n * cb * 8);
  unsigned char* flags = ws;
  ws += align256((size_t)n);
  float* kth = (float*)ws;
  int* nkept = (int*)(ws + 16);
  cudaMemsetAsync(flags, 0, (size_t)n, st);
  nms_gather_kernel<<<cdiv(n, 256), 256, 0, st>>>(boxes, scores, labels, (const long long*)order, (int)n, sorted, nullptr, 0);
  nms_mask_kernel<<<dim3(cb, cb), 64, 0, st>>>((int)n, thresh, sorted, mask, nullptr, 0);
  nms_scan_kernel<<<1, 256, 0, st>>>((int)n, mask, (const long long*)order, sorted, (int)max_det, flags, kth, nkept, nullptr, 0);
  nms_compact_kernel<<<1, 1024, 0, st>>>((int)n, flags, scores, (int)max_det, kth, nkept, (long long*)keep_out, num_keep, nullptr, 0);
  return check_launch("ml_nms");
}

// Batched variant: B images, rows of n_max candidates, per-image counts on the DEVICE (no host sync).
extern "C" int64_t mqdet_ml_nms_batched_workspace_bytes(int64_t B, int64_t n_max) {
  const size_t cb = (size_t)(n_max + 63) / 64;
  return (int64_t)(B * (align256((size_t)n_max * 8) + align256((size_t)n_max * 6 * 4) + align256((size_t)n_max * cb * 8) +
                        align256((size_t)n_max)) + align256((size_t)B * 8) + 256);
}

extern "C" int mqdet_ml_nms_batched(const float* boxes, const float* scores, const float* labels, const int32_t* counts_dev,
                                    int64_t B, int64_t n_max, float thresh, int64_t max_det, int64_t* keep_out,
                                    int32_t* num_keep, void* workspace, void* stream) {
  MQ_REQUIRE(boxes && scores && labels && counts_dev && keep_out && num_keep && workspace, "ml_nms_batched: null pointer");
  MQ_REQUIRE(B > 0 && n_max > 0 && n_max <= NMS_MAX_N, "ml_nms_batched: bad sizes B=%ld n_max=%ld", (long)B, (long)n_max);
  cudaStream_t st = (cudaStream_t)stream;
  const int cb = (int)((n_max + 63) / 64);
  uint8_t* ws = (uint8_t*)workspace;
  long long* order = (long long*)ws;          ws += B * align256((size_t)n_max * 8);
  float* sorted = (float*)ws;                 ws += B * align256((size_t)n_max * 6 * 4);
  unsigned long long* mask = (unsigned long long*)ws;  ws += B * align256((size_t)n_max * cb * 8);
  unsigned char* flags = ws;                  ws += B * align256((size_t)n_max);
  float* kth = (float*)ws;
  int* nkept = (int*)(ws + B * 4);
  // the per-image strides inside the kernels are n_max elements (not the 256-aligned sizes): keep them consistent
  MQ_REQUIRE(align256((size_t)n_max * 8) == (size_t)n_max * 8 && align256((size_t)n_max) == (size_t)n_max,
             "ml_nms_batched: n_max must be a multiple of 256 (got %ld)", (long)n_max);
  cudaMemsetAsync(flags, 0, (size_t)B * n_max, st);
  int np2 = 1;
  while (np2 < n_max) np2 <<= 1;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&argsort_desc_kernel), NMS_MAX_N * 8)) return rc;
  argsort_desc_kernel<<<(unsigned)B, 1024, (size_t)np2 * 8, st>>>(scores, 0, order, counts_dev, (int)n_max);
  nms_gather_kernel<<<dim3(cdiv(n_max, 256), (unsigned)B), 256, 0, st>>>(boxes, scores, labels, order, 0, sorted, counts_dev,
                                                                         (int)n_max);
  nms_mask_kernel<<<dim3(cb, cb, (unsigned)B), 64, 0, st>>>(0, thresh, sorted, mask, counts_dev, (int)n_max);
  nms_scan_kernel<<<(unsigned)B, 256, 0, st>>>(0, mask, order, sorted, (int)max_det, flags, kth, nkept, counts_dev, (int)n_max);
  nms_compact_kernel<<<(unsigned)B, 1024, 0, st>>>(0, flags, scores, (int)max_det, kth, nkept, (long long*)keep_out, num_keep,
                                                   counts_dev, (int)n_max);
  return check_launch("ml_nms_batched");
}
